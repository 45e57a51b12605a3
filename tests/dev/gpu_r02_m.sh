#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02m_pytest.log
tail -5 gpurun_out/r02m_pytest.log
