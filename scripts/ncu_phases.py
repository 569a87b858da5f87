"""Attribute the warp-state samples of an `ncu --page source --print-source cuda,sass --csv` dump of ipm_coop_kernel to
kernel phases.  Samples on lines of CUDA headers (barriers, shuffles) are given to the phase of the nearest preceding
SASS instruction (by address) that maps to repo source."""
import csv, sys, collections

rows = list(csv.reader(open(sys.argv[1])))
sec_file, line = None, None
sass = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        sec_file = r[1].split('/')[-1]
        continue
    if r[0] in ("Function Name", "Line No") or len(r) < 8:
        continue
    if r[0]:
        try:
            line = int(r[0])
        except ValueError:
            pass
        continue
    if not r[2].startswith("0x"):
        continue
    smp = int(r[6]) if r[6] not in ("", "-") else 0
    sass.append((int(r[2], 16), sec_file, line, smp, r[3].strip()))
sass.sort()
RANGES = [(40, 115, 'J\'SJ rank-16 updates (syrk_dmma)'), (224, 256, 'back-substitution'), (380, 450, 'setup_child'),
          (451, 505, 'finish_init'), (506, 545, 'write_results'), (546, 575, 'top: rows'), (576, 625, 'top: J\'SJ staging + J\' products'),
          (626, 725, 'top: variable sweep + tests'), (726, 765, 'A: publish'), (766, 818, 'B: W pass'), (819, 850, 'C: add W, fix'),
          (851, 945, 'C: direction (dots, ftb, mu)'), (946, 1016, 'D: U pass'), (1017, 1160, 'E: line search + update')]


def region(f, l):
    if f == 'bnlp_dense.cuh':
        return 'Cholesky (chol3, bnlp_dense.cuh)'
    if f == 'bnlp_ipm.cuh':
        return 'block_reduce' if 100 <= l <= 145 else 'bnlp_ipm.cuh helpers (dots_grad, tri)'
    if f == 'bnlp_ipm_coop.cuh':
        for lo, hi, name in RANGES:
            if lo <= l <= hi:
                return name
        return 'coop.cuh:%d' % (l // 50 * 50)
    return None


tot = sum(s[3] for s in sass)
reg = collections.Counter()
bar = collections.Counter()
last = '?'
for a, f, l, smp, txt in sass:
    r_ = region(f, l)
    if r_ is not None:
        last = r_
    reg[last] += smp
    if txt.startswith("BAR") or "BAR.SYNC" in txt:
        bar[last] += smp
print("total samples", tot)
for k, v in reg.most_common(30):
    print(f"{100 * v / tot:6.2f}%  (barrier {100 * bar[k] / tot:5.2f}%)  {k}")
