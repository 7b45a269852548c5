"""Generate csrc/smcb_math_tables.inc (run once, output committed): the tables and short polynomials
of the table-assisted exp / log variants of smcb_math.cuh (SMCB_TABLE_MATH=1).

    python particles_b200/csrc/gen_tables.py > particles_b200/csrc/smcb_math_tables.inc

exp:  x = (k / 1024) ln 2 + r,  |r| <= ln2 / 2048,  exp(x) = 2^(k >> 10) * T[k & 1023] * P4(r)   (SMCB_TABLE_MATH=1)
      x = (k / 128) ln 2 + r,   |r| <= ln2 / 256,   exp(x) = 2^(k >> 7) * T[k & 127] * P5(r)     (SMCB_TABLE_MATH=2:
      1 KB table = 8 cache lines, so a divergent lookup touches at most 8 lines)
log:  m in [sqrt(1/2), sqrt(2)) -> interval j (top mantissa bits), r = m * inv_c[j] - 1, |r| <= 2^-7,
      log(m) = nlog_c[j] + log1p(r),  nlog_c[j] = -log(inv_c[j]) for the ROUNDED inv_c[j]
"""
import mpmath as mp

mp.mp.dps = 60


def fit(f, a, b, deg):
    poly, err = mp.chebyfit(f, [a, b], deg + 1, error=True)
    return poly[::-1], err


def relerr(f, coeffs, a, b, n=2001):
    worst = mp.mpf(0)
    for i in range(n):
        x = a + (b - a) * mp.mpf(i) / (n - 1)
        p = sum(c * x ** k for k, c in enumerate(coeffs))
        fx = f(x)
        if fx != 0:
            worst = max(worst, abs(p - fx) / abs(fx))
    return worst


def emit_poly(name, coeffs, note):
    print(f"// {note}")
    print(f"static __constant__ double {name}[{len(coeffs)}] = {{")
    for c in coeffs:
        print(f"    {mp.nstr(mp.mpf(float(c)), 17)},   // {float(c).hex()}")
    print("};")


ln2 = mp.log(2)
h = ln2 / 2048
c, _ = fit(mp.exp, -h, h, 4)
c[0], c[1] = mp.mpf(1), mp.mpf(1)
e = relerr(mp.exp, c, -h, h)
emit_poly("kExp4C", c, f"exp(r) on |r| <= ln2/2048, degree 4 (c0 = c1 = 1), max rel err {mp.nstr(e, 3)}")

print("// T[j] = 2^(j/1024), correctly rounded")
print("static __device__ const double kExp2Tab[1024] = {")
for j in range(1024):
    v = float(mp.power(2, mp.mpf(j) / 1024))
    print(f"    {v!r},")
print("};")

h = ln2 / 256
c, _ = fit(mp.exp, -h, h, 5)
c[0], c[1] = mp.mpf(1), mp.mpf(1)
e = relerr(mp.exp, c, -h, h)
emit_poly("kExp5C", c, f"exp(r) on |r| <= ln2/256, degree 5 (c0 = c1 = 1), max rel err {mp.nstr(e, 3)}")
print("// T[j] = 2^(j/128), correctly rounded")
print("static __device__ const double kExp2Tab128[128] = {")
for j in range(128):
    v = float(mp.power(2, mp.mpf(j) / 128))
    print(f"    {v!r},")
print("};")

# log1p(r) = r + r^2 Q(r), |r| <= 2^-7
fq = lambda r: mp.mpf(-0.5) if r == 0 else (mp.log1p(r) - r) / (r * r)
a = mp.mpf(2) ** -7
c, _ = fit(fq, -a, a, 6)
fl = lambda r: mp.log1p(r)
worst = mp.mpf(0)
for i in range(2001):
    r = -a + 2 * a * mp.mpf(i) / 2000
    if r == 0:
        continue
    p = r + r * r * sum(ck * r ** k for k, ck in enumerate(c))
    worst = max(worst, abs(p - fl(r)) / abs(fl(r)))
emit_poly("kLog1pC", c, f"(log1p(r) - r)/r^2 on |r| <= 2^-7, degree 6, max rel err of log1p {mp.nstr(worst, 3)}")

print("// j = bits 20..13 of the high word of m in [sqrt(1/2), sqrt(2)): j < 128 covers [1/2, 1) in steps of")
print("// 1/256, j >= 128 covers [1, 2) in steps of 1/128; the two intervals touching 1 use c = 1 exactly.")
print("// entry = {inv_c, -log(inv_c)}")
print("static __device__ const double2 kLogTab[256] = {")
for j in range(256):
    if j in (127, 128):
        inv, nl = 1.0, 0.0
    else:
        cen = (mp.mpf(1) / 2 + (mp.mpf(j) + mp.mpf(1) / 2) / 256) if j < 128 else (1 + (mp.mpf(j - 128) + mp.mpf(1) / 2) / 128)
        inv = float(1 / cen)
        nl = float(-mp.log(mp.mpf(inv)))
    print(f"    {{{inv!r}, {nl!r}}},")
print("};")
