"""Generates modulatedgps_b200/csrc/exp_tab.h: the 64-entry table 2^(j/64) and the range-reduction constants of
exp_tab() (stream_kernels.cu), with 60-digit decimal arithmetic.
    python tools/gen_exp_tab.py > modulatedgps_b200/csrc/exp_tab.h"""
import math
from decimal import Decimal, getcontext

getcontext().prec = 60
ln2 = Decimal(2).ln()
c = math.log(2.0) / 64.0
m, e = math.frexp(c)
hi = math.ldexp(math.floor(m * 2 ** 32), e - 32)
lo = float(ln2 / 64 - Decimal(hi))
tab = [float(Decimal(2) ** (Decimal(j) / 64)) for j in range(64)]
print("// 2^(j/64), j = 0..63, correctly rounded, and the range-reduction constants of exp_tab()")
print("// (generated with 60-digit decimal arithmetic by tools/gen_exp_tab.py)")
print("#pragma once")
print("#define EXP_TAB64_VALUES \\")
print(", \\\n".join("    " + ", ".join(repr(x) for x in tab[i:i + 4]) for i in range(0, 64, 4)))
print(f"#define EXP_TAB_L {64.0 / math.log(2.0)!r}       /* 64 / ln 2 */")
print(f"#define EXP_TAB_C_HI {hi!r}   /* ln 2 / 64, top 32 bits */")
print(f"#define EXP_TAB_C_LO {lo!r}   /* remainder */")
# exp2_tab(): argument already in units of ln2/64 (y = x * 64 / ln 2, produced by the z.x contraction itself):
# 2^(rr/64) - 1 = sum_i EXP2_C_i rr^i, EXP2_C_i = (ln2/64)^i / i!, |rr| <= 1/2
for i in range(1, 6):
    ci = (ln2 / 64) ** i / Decimal(math.factorial(i))
    print(f"#define EXP2_C{i} {float(ci)!r}")
