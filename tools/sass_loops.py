#!/usr/bin/env python3
"""List the backward branches (loops) of a SASS dump and the opcode mix of each loop body.
usage: cuobjdump -sass lib.so | python tools/sass_loops.py [kernel-substring]"""
import re, sys, collections
pat = re.compile(r'^\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);')
want = sys.argv[1] if len(sys.argv) > 1 else None
cur = None; funcs = {}
for line in sys.stdin:
    m = re.search(r'Function : (\S+)', line)
    if m: cur = m.group(1); funcs[cur] = []; continue
    m = pat.match(line)
    if m and cur: funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
FP64 = ('DFMA', 'DMUL', 'DADD', 'DSETP')
for name, ins in funcs.items():
    if want and want not in name: continue
    print("==", name, len(ins), "instructions")
    addr = {a: i for i, (a, _) in enumerate(ins)}
    for i, (a, t) in enumerate(ins):
        m = re.search(r'\bBRA(?:\.\w+)*\s+(?:`\(\S+\)|0x([0-9a-f]+))', t)
        if m and m.group(1):
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr:
                body = ins[addr[tgt]:i + 1]
                c = collections.Counter()
                for _, tt in body:
                    op = re.sub(r'^@!?U?P\w+\s+', '', tt).split()[0].split('.')[0]
                    c[op] += 1
                f = sum(c[o] for o in FP64)
                print(f"  loop 0x{tgt:x}..0x{a:x}: {len(body)} instr, fp64-pipe {f} (DFMA {c['DFMA']} DMUL {c['DMUL']} DADD {c['DADD']} DSETP {c['DSETP']}), MUFU {c['MUFU']}, LDS {c['LDS']}, IMAD {c['IMAD']}, MOV {c['MOV']}, FSEL {c['FSEL']}, SEL {c['SEL']}, ISETP {c['ISETP']}, BRA {c['BRA']}, F2F {c['F2F']}, FFMA {c['FFMA']}")
