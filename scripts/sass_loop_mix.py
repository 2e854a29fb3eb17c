"""Static instruction mix of the innermost loop of a kernel in libselfocc_b200.so (cuobjdump -sass), the check done here
before spending GPU time: python scripts/sass_loop_mix.py <mangled-name-substring> [lib.so]
For kernels whose loop contains a warp-vote fast path the interior (taken) side is counted."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, 'selfocc_b200', 'lib', 'libselfocc_b200.so')
names = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
funcs = [m for m in re.findall(r'Function : (\S+)', names) if sys.argv[1] in m]
for fn in funcs:
    txt = subprocess.run(['cuobjdump', '-sass', '-fun', fn, lib], capture_output=True, text=True).stdout
    rows = [(int(m.group(1), 16), m.group(2)) for m in re.finditer(r'/\*([0-9a-f]{4})\*/\s+(.*?);', txt)]
    back = []
    for a, t in rows:
        if 'BRA' in t:
            x = re.findall(r'0x([0-9a-f]+)', t)
            if x and int(x[-1], 16) < a:
                back.append((int(x[-1], 16), a))
    if not back:
        print(fn, ': no loop'); continue
    # the gather loop = the loop with the most global loads that contains no other loop (ties: the shorter one)
    inner = [p for p in back if not any(q != p and p[0] <= q[0] and q[1] <= p[1] for q in back)]
    n_ldg = lambda p: sum(1 for a, t in rows if p[0] <= a <= p[1] and 'LDG' in t)
    lo, hi = max(inner, key=lambda p: (n_ldg(p), -(p[1] - p[0])))
    body = [(a, t) for a, t in rows if lo <= a <= hi]
    hot = body
    votes = [i for i, (a, t) in enumerate(body) if t.startswith('VOTE')]
    if votes:                                    # count the interior (all lanes inside the volume) side of the warp vote only
        br = next((i for i in range(votes[0], len(body)) if re.match(r'@!?P\d BRA', body[i][1])), None)
        if br is not None:
            tgt = int(re.findall(r'0x([0-9a-f]+)', body[br][1])[-1], 16)
            if body[br][1].startswith('@!'):     # exterior is the branch target: interior falls through to `BRA join`
                j = next((i for i in range(br + 1, len(body)) if re.match(r'BRA 0x', body[i][1])), None)
                if j is not None:
                    join = int(re.findall(r'0x([0-9a-f]+)', body[j][1])[-1], 16)
                    hot = body[:j + 1] + [(a, t) for a, t in body if a >= join]
            else:                                # interior is the branch target
                hot = body[:br + 1] + [(a, t) for a, t in body if a >= tgt]
    c = collections.Counter(re.sub(r'^@!?U?P\d\s+', '', t).split()[0].split('.')[0] for a, t in hot)
    grp = lambda *k: sum(v for n, v in c.items() if n in k)
    print('%s\n  innermost loop 0x%x..0x%x: %d instructions on the hot path (of %d in the loop, %d in the kernel)' % (fn, lo, hi, len(hot), len(body), len(rows)))
    print('  fp32 (FADD/FMUL/FFMA/FMNMX/FSEL/FSETP) %d, MUFU %d, conversions (F2I/I2F/FRND/I2FP) %d, loads %d, stores %d, integer/address %d, shuffles/votes %d, branches %d'
          % (grp('FADD', 'FMUL', 'FFMA', 'FMNMX', 'FSEL', 'FSETP'), grp('MUFU'), grp('F2I', 'I2F', 'FRND', 'I2FP'), grp('LDG', 'LD', 'LDS'),
             grp('STG', 'ST', 'STS'), grp('IMAD', 'IADD3', 'LEA', 'ISETP', 'LOP3', 'SHF', 'VIADD', 'VIMNMX', 'MOV', 'SEL', 'IADD'), grp('SHFL', 'VOTE'), grp('BRA', 'BSSY', 'BSYNC')))
    print('  ' + ', '.join('%s %d' % kv for kv in c.most_common(14)))
