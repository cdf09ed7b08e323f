"""Print the polynomial coefficients used in qunundrum_b200/csrc/qmath.cuh.

Development tool (mpmath); the output is pasted into qmath.cuh as decimal
literals with 20 significant digits.
"""
import mpmath as mp

mp.mp.prec = 300
pi = mp.pi


def emit(name, coeffs, comment):
    print(f"// {name}: {comment}")
    ks = sorted(coeffs.keys(), reverse=True)
    first = True
    for k in ks:
        c = mp.nstr(coeffs[k], 20, min_fixed=0, max_fixed=0)
        if first:
            print(f"  double p = {c};  // k = {k}")
            first = False
        else:
            print(f"  p = fma(p, z, {c});  // k = {k}")
    print()


emit("sinpi_kernel", {k: (-1) ** k * pi ** (2 * k + 1) / mp.factorial(2 * k + 1)
                      for k in range(1, 9)}, "(-1)^k pi^(2k+1)/(2k+1)!, k=1..8 (k=0 is pi)")
emit("cospi_kernel", {k: (-1) ** k * pi ** (2 * k) / mp.factorial(2 * k)
                      for k in range(1, 11)}, "(-1)^k pi^(2k)/(2k)!, k=1..10")
emit("sincpi_small", {k: (-1) ** k * pi ** (2 * k) / mp.factorial(2 * k + 1)
                      for k in range(1, 10)}, "(-1)^k pi^(2k)/(2k+1)!, k=1..9")
emit("one_minus_sinc_2pi", {k: (-1) ** (k + 1) * (2 * pi) ** (2 * k) / mp.factorial(2 * k + 1)
                            for k in range(1, 10)}, "(-1)^(k+1) (2pi)^(2k)/(2k+1)!, k=1..9")
emit("one_minus_ecote", {k: mp.mpf(2) ** (2 * k) * abs(mp.bernoulli(2 * k)) / mp.factorial(2 * k)
                         for k in range(1, 13)}, "2^(2k)|B_2k|/(2k)!, k=1..12")
# sinc(pi u)^2 * pi^2 = (sin(pi u)/u)^2 small-u polynomial: pi * sincpi
emit("pi_sincpi_small", {k: (-1) ** k * pi ** (2 * k + 1) / mp.factorial(2 * k + 1)
                         for k in range(0, 10)}, "(-1)^k pi^(2k+1)/(2k+1)!, k=0..9")
# 1/sinc(z)^2 = 1 + z^2/3 + z^4/15 + 2 z^6/189 + ...
z = mp.mpf(0)
ser = mp.taylor(lambda t: (t / mp.sin(t)) ** 2 if t != 0 else mp.mpf(1), 0, 10)
print("// 1/sinc(z)^2 Taylor:", [mp.nstr(c, 20) for c in ser[::2]])
