"""Timeline of CTA (0,0,0) of the round-2 attention kernel (selftest attn ... <variant> <emu> <trace file>).
Regions 0 / 1: thread 0 of softmax warpgroup A / B -- 1 = S(j) observed, 2 = S in registers, 3 = reference maximum
published, 4 = P(j) complete and released;  region 2: the issuer -- 10/11 = first half of P(j) observed (j even/odd),
12/13 = P.V(j) issued and committed, 14/15 = Q.K^T(j+3) issued."""
import sys

import numpy as np

for name in sys.argv[1:]:
    a = np.fromfile(name, dtype=np.uint64).reshape(3, 4096)
    print("====", name)
    per_wg = []
    for wg in (0, 1):
        r = a[wg][a[wg] != 0]
        ev = [(int(x >> np.uint64(8)), int(x & np.uint64(0xff))) for x in r]
        its, cur = [], {}
        for t, i in ev:
            cur[i] = t
            if i == 4:
                its.append(cur)
                cur = {}
        per_wg.append(its)
        its = its[2:-1]
        if len(its) < 3:
            continue
        period = np.mean([its[k + 1][1] - its[k][1] for k in range(len(its) - 1)])
        f = lambda g: np.mean([g(x) for x in its])
        wait = np.mean([its[k + 1][1] - its[k][4] for k in range(len(its) - 1)])
        print(f"WG {wg}: {len(its)} blocks, period {period:.0f} | tmem ld {f(lambda x: x[2] - x[1]):.0f}  max+token "
              f"{f(lambda x: x[3] - x[2]):.0f}  exp+P store+release {f(lambda x: x[4] - x[3]):.0f} | busy "
              f"{f(lambda x: x[4] - x[1]):.0f} | row sum + wait for next S {wait:.0f}")
    r = a[2][a[2] != 0]
    ev = [(int(x >> np.uint64(8)), int(x & np.uint64(0xff))) for x in r]
    obs = [t for t, i in ev if i in (10, 11)]
    pv = [t for t, i in ev if i in (12, 13)]
    qk = [t for t, i in ev if i in (14, 15)]
    n = min(len(obs), len(pv), len(qk))
    if n > 6:
        sl = slice(3, n - 1)
        d = np.diff(np.array(obs[:n]))[3:-1]
        print(f"issuer: {n} blocks, period {d.mean():.0f} (min {d.min()}, max {d.max()}) | P observed -> P.V committed "
              f"{np.mean(np.array(pv[:n])[sl] - np.array(obs[:n])[sl]):.0f} | -> Q.K^T issued "
              f"{np.mean(np.array(qk[:n])[sl] - np.array(pv[:n])[sl]):.0f} | idle until next P "
              f"{np.mean(np.array(obs[1:n + 1])[3:n - 2] - np.array(qk[:n])[3:n - 2]):.0f}")
    # release -> observation latency
    rel = sorted([x[4] for its in per_wg for x in its])
    m = min(len(rel), len(obs))
    if m > 6:
        print(f"P released (softmax thread 0) -> issuer passes the wait: {np.mean(np.array(obs[:m])[3:-1] - np.array(rel[:m])[3:-1]):.0f}")
