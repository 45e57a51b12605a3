"""masurvival (B200 build): batched drop-in for the step path of
KRLGroup/gym-ma-survival-2d.  `from masurvival.envs import MaSurvivalVec`."""
