"""Independent exact big-integer model of the BFV arithmetic on pplp's path (pure Python ints).

It restates the *mathematics* (ring Z_Q[x]/(x^N+1), CRT, exact rounding), not SEAL's code, so it pins every
RNS routine of the oracle whose result is a uniquely defined function of its inputs.
Small N only (schoolbook products).  TEST INFRASTRUCTURE.
"""
from functools import reduce


def prod(xs):
    return reduce(lambda a, b: a * b, xs, 1)


def crt(residues, moduli):
    Q = prod(moduli)
    x = 0
    for r, q in zip(residues, moduli):
        p = Q // q
        x += int(r) * p * pow(p, -1, q)
    return x % Q


def centre(x, Q):
    x %= Q
    return x - Q if x > Q // 2 else x


def poly_crt(limbs, moduli):
    """limbs: [k][N] residues -> list of N integers mod Q."""
    n = len(limbs[0])
    return [crt([limbs[j][i] for j in range(len(moduli))], moduli) for i in range(n)]


def negacyclic_mul(a, b, Q):
    n = len(a)
    out = [0] * n
    for i, x in enumerate(a):
        if x == 0:
            continue
        for j, y in enumerate(b):
            k = i + j
            if k < n:
                out[k] = (out[k] + x * y) % Q
            else:
                out[k - n] = (out[k - n] - x * y) % Q
    return out


def bitrev(x, bits):
    r = 0
    for i in range(bits):
        r = (r << 1) | ((x >> i) & 1)
    return r


def ntt_definition(a, psi, q):
    """out[j] = sum_i a_i psi^(i*(2*bitrev(j)+1)) — the function SEAL's forward NTT computes (bit-reversed output)."""
    n = len(a)
    bits = n.bit_length() - 1
    out = []
    for j in range(n):
        e = 2 * bitrev(j, bits) + 1
        w = pow(psi, e, q)
        acc, p = 0, 1
        for x in a:
            acc = (acc + int(x) * p) % q
            p = p * w % q
        out.append(acc)
    return out


def round_scale(m, Q, t):
    """round(Q*m/t) with ties up = floor((Q*m + floor((t+1)/2)) / t)."""
    return (Q * m + (t + 1) // 2) // t


def decrypt_exact(c0, c1, s, Q, t):
    """m = round(t * [c0 + c1*s]_Q / Q) mod t, centred representative."""
    n = len(c0)
    cs = negacyclic_mul(c1, s, Q)
    out = []
    for i in range(n):
        v = centre(c0[i] + cs[i], Q)
        num = t * v
        m = (2 * num + Q) // (2 * Q)  # round to nearest (ties up)
        out.append(m % t)
    return out
