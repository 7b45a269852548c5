"""Generate the polynomial coefficients of csrc/smcb_math.cuh (run once, output committed).

    python particles_b200/csrc/gen_coeffs.py > particles_b200/csrc/smcb_math_coeffs.inc

Near-minimax (Chebyshev-interpolation) fits computed with mpmath at 60 digits; the printed
max relative errors are those of the polynomials in exact arithmetic.
"""
import mpmath as mp

mp.mp.dps = 60


def fit(f, a, b, deg):
    poly, err = mp.chebyfit(f, [a, b], deg + 1, error=True)
    return poly[::-1], err            # ascending powers


def relerr(f, coeffs, a, b, n=4001):
    worst = mp.mpf(0)
    for i in range(n):
        x = a + (b - a) * mp.mpf(i) / (n - 1)
        p = sum(c * x ** k for k, c in enumerate(coeffs))
        fx = f(x)
        if fx != 0:
            worst = max(worst, abs(p - fx) / abs(fx))
    return worst


def emit(name, coeffs, note):
    print(f"// {note}")
    print(f"static __constant__ double {name}[{len(coeffs)}] = {{")
    for c in coeffs:
        print(f"    {mp.nstr(mp.mpf(float(c)), 17)},   // {float(c).hex()}")
    print("};")


ln2 = mp.log(2)
# exp(r), |r| <= ln2/2
c, _ = fit(mp.exp, -ln2 / 2, ln2 / 2, 11)
e = relerr(mp.exp, c, -ln2 / 2, ln2 / 2)
emit("kExpC", c, f"exp(r) on |r| <= ln2/2, degree 11, max rel err {mp.nstr(e, 3)}")

# sin(pi r)/r and cos(pi r) as polynomials in z = r^2, |r| <= 1/4
fs = lambda z: mp.pi if z == 0 else mp.sin(mp.pi * mp.sqrt(z)) / mp.sqrt(z)
c, _ = fit(fs, 0, mp.mpf(1) / 16, 6)
e = relerr(fs, c, 0, mp.mpf(1) / 16)
emit("kSinPiC", c, f"sin(pi r)/r in z = r^2 on [0, 1/16], degree 6, max rel err {mp.nstr(e, 3)}")
fc = lambda z: mp.cos(mp.pi * mp.sqrt(z))
c, _ = fit(fc, 0, mp.mpf(1) / 16, 7)
e = relerr(fc, c, 0, mp.mpf(1) / 16)
emit("kCosPiC", c, f"cos(pi r) in z = r^2 on [0, 1/16], degree 7, max rel err {mp.nstr(e, 3)}")

# log(m) = 2 atanh(s), s = (m-1)/(m+1), m in [sqrt(1/2), sqrt(2)]: 2*atanh(s)/s in z = s^2
smax = (mp.sqrt(2) - 1) / (mp.sqrt(2) + 1)
fl = lambda z: mp.mpf(2) if z == 0 else 2 * mp.atanh(mp.sqrt(z)) / mp.sqrt(z)
c, _ = fit(fl, 0, smax ** 2, 7)
e = relerr(fl, c, 0, smax ** 2)
emit("kLogC", c, f"2 atanh(s)/s in z = s^2 on [0, {mp.nstr(smax**2, 6)}], degree 7, max rel err {mp.nstr(e, 3)}")

print(f"// ln2 split: hi has 32 trailing zero bits so that k * ln2_hi is exact for |k| < 2^20")
hi = mp.mpf(float(ln2))
import struct
bits = struct.unpack("<Q", struct.pack("<d", float(ln2)))[0] & ~((1 << 32) - 1)
hi = mp.mpf(struct.unpack("<d", struct.pack("<Q", bits))[0])
lo = ln2 - hi
print(f"#define SMCB_LN2_HI {mp.nstr(hi, 17)}")
print(f"#define SMCB_LN2_LO {mp.nstr(mp.mpf(float(lo)), 17)}")
print(f"#define SMCB_LOG2E {mp.nstr(mp.mpf(float(1 / ln2)), 17)}")
