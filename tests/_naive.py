"""Pure-Python big-int restatements used to pin the C++ oracle on small cases.

Independent of both the oracle and the CUDA path: plain O(n^2) field sums."""
P = 2**64 - 2**32 + 1
GENERATOR = 7
ROOT32 = 1753635133440165772  # p3-goldilocks two-adic generator of order 2^32
W_EXT = 7


def two_adic_generator(bits):
    return pow(ROOT32, 1 << (32 - bits), P)


def rev(x, bits):
    r = 0
    for i in range(bits):
        r = (r << 1) | ((x >> i) & 1)
    return r


def dft(col):
    """dft(f)_k = sum_j f_j w^{jk}."""
    n = len(col)
    if n == 1:
        return list(col)
    w = two_adic_generator(n.bit_length() - 1)
    return [sum(col[j] * pow(w, j * k, P) for j in range(n)) % P for k in range(n)]


def idft(col):
    n = len(col)
    f = dft(col)
    ninv = pow(n, P - 2, P)
    return [f[(n - j) % n] * ninv % P for j in range(n)]


def poly_eval(coeffs, x):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % P
    return acc


def coset_lde_bitrev(col, added_bits, shift=GENERATOR):
    """stored[i] = P(shift * w_{nB}^{rev(i)})."""
    n = len(col)
    coeffs = idft(col)
    big = n << added_bits
    lb = big.bit_length() - 1
    w = two_adic_generator(lb)
    return [poly_eval(coeffs, shift * pow(w, rev(i, lb), P) % P) for i in range(big)]


# --- extension field F_p[X]/(X^2 - 7) ---------------------------------------------------------
def e_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def e_sub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def e_mul(a, b):
    return ((a[0] * b[0] + W_EXT * a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def e_inv(a):
    norm = (a[0] * a[0] - W_EXT * a[1] * a[1]) % P
    ni = pow(norm, P - 2, P)
    return (a[0] * ni % P, (-a[1]) * ni % P)


def e_pow(a, e):
    acc = (1, 0)
    while e:
        if e & 1:
            acc = e_mul(acc, a)
        a = e_mul(a, a)
        e >>= 1
    return acc


def poly_eval_ext(coeffs, z):
    acc = (0, 0)
    for c in reversed(coeffs):
        acc = e_mul(acc, z)
        acc = ((acc[0] + c) % P, acc[1])
    return acc
