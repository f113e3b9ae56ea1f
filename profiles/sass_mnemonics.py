'''Count the SASS mnemonics that prove the Blackwell paths (tcgen05.mma = UTCHMMA / UTCQMMA, TMA = UTMALDG / UTMASTG, TMEM =
LDTM / STTM, mbarrier = SYNCS, tcgen05.commit = UTCBAR, redux = CREDUX) per kernel of the built library (no GPU needed):

    python profiles/sass_mnemonics.py > profiles/r02/sass_mnemonics.txt
'''
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ('UTCHMMA', 'UTCQMMA', 'UTMALDG', 'UTMASTG', 'LDTM', 'STTM', 'SYNCS', 'UTCBAR', 'CREDUX', 'MUFU.EX2', 'MEMBAR.ALL.GPU')


def main():
    so = os.path.join(ROOT, 'flexdiffuse_b200', 'libflexdiffuse_b200.so')
    sass = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
    counts, order, cur = {}, [], None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r'\(anonymous namespace\)::', '', cur).split('(')[0]
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur:
            for w in WANT:
                if re.search(r'\b' + re.escape(w) + r'\b', line):
                    counts[cur][w] += 1
    head = subprocess.run(['git', '-C', ROOT, 'rev-parse', '--short', 'HEAD'], capture_output=True, text=True).stdout.strip()
    print(f'# SASS mnemonic counts per kernel of libflexdiffuse_b200.so (cuobjdump -sass, built from {head}); tcgen05.mma = UTCHMMA, '
          'TMA = UTMALDG/UTMASTG, TMEM = LDTM/STTM, tcgen05.commit = UTCBAR, mbarrier = SYNCS')
    for k in sorted(order):
        print(f'{k}: ' + ', '.join(f'{w}={n}' for w, n in sorted(counts[k].items())))


if __name__ == '__main__':
    main()
