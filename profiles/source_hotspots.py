#!/usr/bin/env python
"""Where k_step's warp time goes, by source function: joins the per-instruction stall samples of an
`ncu --set full --import-source on` report with the line table of the same build's SASS.

    ncu -i gpurun_out/r02f_k_step_2v2.ncu-rep --page source --csv > /tmp/src.csv
    cuobjdump -xelf all gym-ma-survival-2d_b200/csrc/build/msv_kernels.o        # -> msv_kernels.sm_100a.cubin
    nvdisasm -g msv_kernels.sm_100a.cubin > /tmp/all.sass
    python profiles/source_hotspots.py /tmp/src.csv /tmp/all.sass '_Z6k_stepILi4ELi4ELi4ELi4E' [--lines]

The report's rows and the disassembly list the same instructions in the same order (checked by opcode).  A
sample is attributed to the innermost inlined function of its line; the out-of-line (cold) functions are
lumped under their own names in brackets."""
import bisect
import collections
import csv
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = os.path.join(ROOT, 'gym-ma-survival-2d_b200', 'csrc')


def funcs(path):
    out = []
    for i, l in enumerate(open(path).read().split('\n'), 1):
        m = re.match(r'^\s{0,2}(?:template\s*<[^>]*>\s*)?(?:DEV|COLD\d|__device__|static|__global__|inline)\b.*?([A-Za-z_][A-Za-z_0-9]*)\s*\(', l)
        if m and not l.strip().startswith('//') and ';' not in l.split('{')[0]:
            out.append((i, m.group(1)))
    return out


def main():
    src_csv, sass, kernel = sys.argv[1:4]
    F = {f: funcs(os.path.join(BASE, f)) for f in ('msv_env.cuh', 'msv_device.cuh', 'msv_kernels.cu')}

    def fn(file, line):
        if file not in F:
            return file
        a = F[file]
        k = bisect.bisect_right([x[0] for x in a], line) - 1
        return file.split('.')[0][4:] + ':' + (a[k][1] if k >= 0 else '?')

    ins, cur, region, on = [], None, 'MAIN', False
    for l in open(sass):
        if l.startswith('.text.'):
            on = kernel in l
            region = 'MAIN'
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
            continue
        m = re.search(r'\.type\s+\$[^$]+\$([^,]+),@function', l) or re.search(r'\.type\s+\$(__[^,]+),@function', l)
        if m:
            region = m.group(1)
            continue
        m = re.match(r'^\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
        if m:
            ins.append((m.group(2).strip(), cur, region))
    R = list(csv.reader(open(src_csv)))
    hdr, rows = R[1], R[2:]
    assert len(rows) == len(ins), (len(rows), len(ins))
    ci = {n: i for i, n in enumerate(hdr)}
    S, IE, TI = ci['# Samples'], ci['Instructions Executed'], ci['Thread Instructions Executed']
    stalls = [n for n in hdr if n.startswith('stall_') and '(' not in n]
    tot = sum(int(r[S]) for r in rows)
    totie = sum(int(r[IE]) for r in rows)
    print('kernel', R[0][1])
    print('SASS instructions', len(ins), '| samples', tot, '| warp instructions executed', totie, '| active lanes per instruction %.2f' % (sum(int(r[TI]) for r in rows) / totie))
    st = collections.Counter()
    for r in rows:
        for n in stalls:
            st[n] += int(r[ci[n]])
    print('stall reasons (share of samples):', ', '.join('%s %.1f%%' % (k[6:], 100 * v / tot) for k, v in st.most_common(10)))

    def agg(keyf, title, n):
        c, ie, ti, sz, sb = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
        for r, i in zip(rows, ins):
            k = keyf(i)
            c[k] += int(r[S]); ie[k] += int(r[IE]); ti[k] += int(r[TI]); sz[k] += 1
            for nme in stalls:
                sb[k][nme] += int(r[ci[nme]])
        print('\n== ' + title)
        print('%7s %6s %9s %6s %6s  %-36s %s' % ('samples', '%', 'warp inst', 'lanes', 'SASS', 'where', 'top stall reasons'))
        for k, v in c.most_common(n):
            top = ', '.join('%s %d%%' % (a[6:], 100 * b / max(v, 1)) for a, b in sb[k].most_common(3))
            print('%7d %5.1f%% %9d %6.1f %6d  %-36s %s' % (v, 100 * v / tot, ie[k], ti[k] / max(ie[k], 1), sz[k], str(k)[:36], top))

    agg(lambda i: re.sub(r'^_ZN3EnvILi\dELi\dELi\d+ELi\dEE\d+', '', i[2])[:36], 'by SASS function (MAIN = the kernel body with everything inlined into it)', 24)
    agg(lambda i: (fn(*i[1]) if i[1] else '?') if i[2] == 'MAIN' else '[' + re.sub(r'^_ZN3EnvILi\dELi\dELi\d+ELi\dEE\d+', '', i[2])[:20] + ']',
        'by innermost source function (samples at a barrier are charged to the code right after it)', 60)
    if '--lines' in sys.argv:
        agg(lambda i: (i[1][0][4:9] + ':%d' % i[1][1]) if i[1] and i[2] == 'MAIN' else '-', 'kernel body by source line', 80)


if __name__ == '__main__':
    main()
