"""Range proof of the FP64-pipe transforms (pplp_b200/csrc/modarith.cuh mulmod_f64, ntt.cuh L = 3 / 4, ntt32.cuh), on the CPU.

The kernels keep residues as exact integers in doubles; that is only sound while every multiplicand stays within 2^51
(so that the quotient estimate rounds to an integer) and every intermediate within 2^53.  Random GPU parity tests seldom
reach the worst case, so this file (a) replays the arithmetic of one product with exact rational fused multiply-adds on
adversarial operands and (b) propagates worst-case magnitudes through the stage / pass / reduction schedule of every
kernel variant for the widest modulus each variant accepts.  The schedules below restate the kernels' control flow; the
constants (44 / 49 bits, the 96 q rule, reductions after every second inverse stage) are the ones in the headers."""
import random
from fractions import Fraction

import pytest

TWO51, TWO52, TWO53 = 1 << 51, 1 << 52, 1 << 53
MAGIC = 1.5 * TWO52


def fma(a, b, c):
    return float(Fraction(a) * Fraction(b) + Fraction(c))   # one rounding, to nearest even


def mulmod_f64(a, w, wi, q):
    h = a * w
    lo = fma(a, w, -h)
    c = fma(a, wi, MAGIC) - MAGIC
    return (fma(-c, q, h) + lo), c


def reduce_sym_f64(a, qi, q):
    c = fma(a, qi, MAGIC) - MAGIC
    return fma(-c, q, a)


@pytest.mark.parametrize("bits,amax", [(44, TWO51), (43, TWO51), (49, TWO51), (45, TWO51), (36, TWO51), (27, 1 << 40)])
def test_one_product_is_exact_up_to_the_budget(bits, amax):
    rng = random.Random(bits)
    q = (1 << bits) - rng.randrange(1, 1 << 12) * 2 - 1
    qd, qi = float(q), 1.0 / float(q)
    cases = [(amax, q - 1), (-amax, q - 1), (amax - 1, q - 1), (amax, 1), (-amax, (q + 1) // 2), (amax, (q - 1) // 2), (0, q - 1), (1, q - 1), (-1, 1)]
    cases += [(rng.randrange(-amax, amax + 1), rng.randrange(0, q)) for _ in range(3000)]
    cases += [(s * (amax - rng.randrange(0, 1 << 20)), q - 1 - rng.randrange(0, 1 << 10)) for s in (1, -1) for _ in range(500)]
    for a, w in cases:
        wi = w / q                                   # correctly rounded: both are exact doubles
        t, c = mulmod_f64(float(a), float(w), wi, qd)
        assert c == int(c) and t == int(t)
        assert int(t) == a * w - int(c) * q          # the product is EXACT
        assert abs(int(t)) <= q * (0.5 + abs(a) * 2.0 ** -53) + 1
        r = reduce_sym_f64(float(a), qi, qd)
        assert r == int(r) and (int(r) - a) % q == 0 and abs(int(r)) <= q * (0.5 + abs(a) * 2.0 ** -53) + 1


# ---- worst-case magnitude propagation (units of q) -------------------------------------------------------------------
class Budget:
    def __init__(self, bits):
        self.q = (1 << bits) - 1                     # the widest modulus of that many bits
        self.limit = Fraction(TWO51, self.q)         # multiplicand budget in units of q
        self.worst = Fraction(0)

    def product(self, m):                            # bound of a*w mod q for |a| <= m q
        assert m <= self.limit, f"multiplicand {float(m):.3f} q exceeds 2^51 = {float(self.limit):.3f} q"
        self.worst = max(self.worst, m)
        return Fraction(1, 2) + m * self.q / TWO53

    def reduced(self, m):
        assert m * self.q < TWO53
        return Fraction(1, 2) + m * self.q / TWO53


def ct_pass(b, R, B):        # Cooley-Tukey, in-thread radix-2^R group: x' = x + t(y), y' = x - t(y)
    n = 1 << R
    for bit in reversed(range(R)):
        for r in range(n):
            if not r & (1 << bit):
                t = B.product(b[r | 1 << bit])
                b[r] = b[r | 1 << bit] = b[r] + t
    assert max(b) * B.q < TWO53
    return b


def gs_stage(b, bit, B, fold=False):   # Gentleman-Sande: x' = x + y, y' = (x - y) w;  fold: both outputs are products
    for r in range(len(b)):
        if not r & (1 << bit):
            s = b[r] + b[r | 1 << bit]
            assert s * B.q < TWO53
            b[r | 1 << bit] = B.product(s)
            b[r] = B.product(s) if fold else s


def shape16(logm):
    r0 = 4 if logm % 4 == 0 else logm % 4
    return r0, (logm - r0) // 4


@pytest.mark.parametrize("logm", [10, 11, 12, 13, 14])
@pytest.mark.parametrize("level,bits", [(3, 44), (4, 49)])
def test_16_per_thread_schedules_stay_within_budget(logm, level, bits):
    if level == 4 and logm < 12:
        pytest.skip("L = 4 is instantiated for N >= 4096")
    B = Budget(bits)
    r0, nfull = shape16(logm)
    # forward (block_ntt_forward): coarse pass, then nfull radix-16 passes; L = 4 reduces before every full pass
    x = Fraction(4 if level == 3 else 1)            # L = 3 accepts inputs below 4q, L = 4 canonical inputs
    x = max(ct_pass([x] * (1 << r0), r0, B))
    for _ in range(nfull):
        if level == 4:
            x = B.reduced(x)
        x = max(ct_pass([x] * 16, 4, B))
    assert B.reduced(x) <= 1                         # forward_canon's reduction brings it to [-q/2, q/2]
    if level == 3:
        assert x <= 16                               # forward_lazy shifts by 16 q
    else:
        assert x <= 4                                # ... by 4 q
    # inverse (block_ntt_inverse): nfull radix-16 passes (stages V = 3..0), then the coarse pass with N^-1 folded in
    x = Fraction(2) if level == 3 else Fraction(1)   # inputs below 2q; L = 4 centres them to (-q, q) while converting
    for _ in range(nfull):
        b = [x] * 16
        for k, bit in enumerate(range(4)):           # executed order: gap 1 first = register bit 0
            gs_stage(b, bit, B)
            if level == 4 and k in (1, 3):           # reduce_sums after the 2nd and the 4th executed stage
                for r in range(16):
                    if not r & (1 << bit):
                        b[r] = B.reduced(b[r])
        if level == 3:
            b[0] = B.reduced(b[0])                   # the all-sums register
        x = max(b)
    b = [x] * (1 << r0)
    for k, bit in enumerate(range(r0)):
        last = k == r0 - 1
        gs_stage(b, bit, B, fold=last)
        if level == 4 and not last and ((r0 == 4 and k == 1) or (r0 == 3 and k == 1)):
            for r in range(1 << r0):
                if not r & (1 << bit):
                    b[r] = B.reduced(b[r])
    assert max(b) <= 1                               # exit: products, shifted by q into (0, 2q)


def gs_bound(j, R, X):       # ntt32.cuh gs_bound
    if j == 0:
        return X * (1 << R)
    msb = max(b for b in range(R) if j >> b & 1)
    return 0.75 * (1 << (R - 1 - msb))


@pytest.mark.parametrize("logm", [11, 12, 13])
def test_32_per_thread_schedule_stays_within_budget(logm):
    B = Budget(44)
    sb = logm - 10
    # forward: A (5 stages), B (sb stages on the low sb bits of the register index), C (5 stages); inputs below 4q
    x = max(ct_pass([Fraction(4)] * 32, 5, B))
    x = max(ct_pass([x] * (1 << sb), sb, B))
    x = max(ct_pass([x] * 32, 5, B))
    assert B.reduced(x) <= 1 and x <= 16
    # inverse: C' (5), reduce by the 96 q rule, B' (sb), reduce, A' (4 plain stages + the folded one)
    b = [Fraction(2)] * 32
    for bit in range(5):
        gs_stage(b, bit, B)
    for e in range(32):
        assert b[e] <= gs_bound(e, 5, 2.0) + 1e-9    # the closed form the kernel's rule is built on
        if gs_bound(e, 5, 2.0) * (1 << sb) > 96:
            b[e] = B.reduced(b[e])
    assert max(b) <= 12
    groups = []
    for hi in range(32 >> sb):                       # pass B' acts on groups of 2^sb registers
        g = [b[(hi << sb) | lo] for lo in range(1 << sb)]
        for bit in range(sb):
            gs_stage(g, bit, B)
        for lo in range(1 << sb):
            if gs_bound(lo, sb, 12.0) * 32 > 96:
                g[lo] = B.reduced(g[lo])
        groups += g
    x = max(groups)
    b = [x] * 32
    for bit in range(5):
        gs_stage(b, bit, B, fold=(bit == 4))
    assert max(b) <= 1
    assert B.worst <= 96 + 1e-6                      # the margin the headers quote: 96 q of 128 q


def test_32_per_thread_wide_schedule_stays_within_budget():
    """ntt32.cuh WIDE rule set (N = 16384: 5 + 4 + 5 stages, moduli up to 49 bits, budget 4 q)."""
    B = Budget(49)
    # forward: canonical inputs; passes B and C start with a reduction of all registers
    x = max(ct_pass([Fraction(1)] * 32, 5, B))
    assert x <= Fraction(19, 4)
    x = max(ct_pass([B.reduced(x)] * 16, 4, B))
    x = max(ct_pass([B.reduced(x)] * 32, 5, B))
    assert x * B.q < TWO53 and B.reduced(x) <= 1                 # ntt32_canon reduces, then shifts by q
    # inverse: inputs centred to (-q, q); "a" stage, then "b" stage whose sum outputs are reduced, 14 stages, the last one folded
    def stage(b, bit, reduce_sums, fold=False):
        gs_stage(b, bit, B, fold=fold)
        if reduce_sums:
            for r in range(len(b)):
                if not r & (1 << bit):
                    b[r] = B.reduced(b[r])
    b = [Fraction(1)] * 32
    for k, bit in enumerate(range(5)):                            # C': global stages 1..5 = a b a b a
        stage(b, bit, reduce_sums=(k % 2 == 1))
    assert max(b) <= 2
    g = [max(b)] * 16
    for k, bit in enumerate(range(4)):                            # B': 6..9 = b a b a
        stage(g, bit, reduce_sums=(k % 2 == 0))
    assert max(g) <= 2
    b = [max(g)] * 32
    for k, bit in enumerate(range(5)):                            # A': 10..14 = b a b a b(folded)
        stage(b, bit, reduce_sums=(k % 2 == 0 and k != 4), fold=(k == 4))
    assert max(b) <= 1                                            # products, shifted by q into (0, 2q)
    assert B.worst <= B.limit
    # relinearisation stage 2 feeds the inverse with reduced sums of products (<= 0.57 q + ...): within the "a" stage's 1 q
    assert B.reduced(Fraction(8)) <= 1


def test_budget_model_rejects_a_modulus_that_is_too_wide():
    B = Budget(45)                                   # ntt32 / L = 3 are limited to 44 bits for this reason
    with pytest.raises(AssertionError):
        b = [Fraction(2)] * 32
        for bit in range(5):
            gs_stage(b, bit, B)
        b[0] = B.reduced(b[0])
        g = [max(b)] * 8
        for bit in range(3):
            gs_stage(g, bit, B)
