"""Small text-path workload for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import qunundrum_b200 as qb
from oracle import text as ot
from tests.test_text_format import value_sets, adversarial, decimal_tokens, midpoint_tokens, join

ctx = qb.Context(0)
v, gold = adversarial()
assert ctx.text_format(v) == gold
back, used = ctx.text_parse(gold, v.size)
assert used == len(gold)
for n in (0, 1, 255, 256, 257, 1023, 1024, 1025, 5000):
    vals = value_sets(1 + n, max(n, 1))["full_range"][:n]
    t = ctx.text_format(vals, np.longdouble(3))
    assert t == ot.format_ld24(np.append(vals, np.longdouble(3)))
    b, used = ctx.text_parse(t, n + 1)
    assert used == len(t)
toks = decimal_tokens(3, 3000) + midpoint_tokens(4, 500)
t = join(toks, b" \n")
got, used = ctx.text_parse(t, len(toks))
assert used == len(t)
d = np.random.default_rng(1).standard_normal(3000)
assert ctx.text_format(d) == ot.format_ld24(d.astype(np.longdouble))
# a little of the integrators too
P = qb.Parameters(256, 2, 2 ** 255 + 12345, 2 ** 256 - 98765)
ctx.slice2d_batch(P, 0, True, 32, [257, -258], [256, 257])
ctx.slice1d_batch(P, 0, True, 64, [250, -251])
print("sanitize workload ok")
