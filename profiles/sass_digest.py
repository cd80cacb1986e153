#!/usr/bin/env python
"""Instruction histogram of every kernel in a generated solver library (cuobjdump -sass), the digest VERDICT r1 asked for:
   python profiles/sass_digest.py cvxpygen_b200/_generated/mpc_12_4_10/libcpg_b200.so > profiles/r2_sass_mpc_12_4_10.md
Columns: total SASS instructions and the opcodes that characterise the design -- UBLKCP (TMA bulk copy), SYNCS (mbarrier),
DFMA / DMMA / DADD / DMUL (FP64 pipe), LDS / STS (shared memory), LDL / STL (register spills to local memory), LDG / STG,
SHFL, BAR, ATOMS / RED."""
import collections, re, subprocess, sys

OPS = ['UBLKCP', 'SYNCS', 'DFMA', 'DMMA', 'DADD', 'DMUL', 'MUFU', 'LDS', 'STS', 'LDL', 'STL', 'LDG', 'STG', 'LDC', 'SHFL', 'BAR', 'ATOMS', 'ATOMG', 'RED', 'REDUX']


def digest(path):
    out = subprocess.run(['cuobjdump', '-sass', path], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = m.group(1); kernels[cur] = collections.Counter(); continue
        m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', line)
        if m and cur:
            op = m.group(1)
            kernels[cur]['_total'] += 1
            kernels[cur][op.split('.')[0]] += 1
    return kernels


def demangle(name):
    try:
        return subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip() or name
    except Exception:
        return name


if __name__ == '__main__':
    for path in sys.argv[1:]:
        print(f'# SASS digest of `{path}` (sm_100a)\n')
        print('| kernel | total | ' + ' | '.join(OPS) + ' |')
        print('|---|---|' + '---|' * len(OPS))
        for k, c in digest(path).items():
            nm = re.sub(r'\(.*', '', demangle(k)).replace('void ', '').replace('cpgb200::', '').replace('<(anonymous namespace)::Fam', '<Fam')
            print(f'| `{nm[:60]}` | {c["_total"]} | ' + ' | '.join(str(c.get(o, 0)) for o in OPS) + ' |')
        print()
