"""Design aid (CPU, oracle): per tree level of a BASELINE configs[1]-shaped partition, how the open nodes' sizes are distributed and at which level
every open node of a tree fits a given cap -- what sizes tree_tail_kernel's hand-over rule."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import pyoracle as O

rng = np.random.default_rng(5)
P, fp, ratio, m, T, mtry = (800, 2, 3, 26, 10, 5) if len(sys.argv) < 2 else (360, 2, 3, 26, 10, 5)
nPos, nNeg = (fp + 1) * P, (fp + 1) * P * ratio
Ntr = nPos + nNeg
base = rng.standard_normal((P, m)); base[:, :5] += 1.0
pos = [base]
for _ in range(fp):                      # SMOTE-like: convex combinations of a positive and one of its neighbours (here: random other positive nearby)
    j = rng.integers(0, P, P); a = rng.random((P, 1))
    pos.append(base + a * (base[j] - base) * 0.3)
X = np.vstack(pos + [rng.standard_normal((nNeg, m))])
lab = np.r_[np.ones(nPos), 2 * np.ones(nNeg)]
D = np.ascontiguousarray(np.vstack([X.T, lab[None, :]]))
seed = 12345
F = O.o_forest_train(D, T, mtry, seed)
caps = [32, 64, 128, 256, 512]
for t in range(T):
    tr = F.tree(t)
    tree_seed = (seed * (t + 1)) & 0xFFFFFFFF
    boot, _ = O.o_tree_draws(tree_seed, Ntr, m, mtry, 1)
    rows = np.unique(boot.astype(np.int64))
    node = np.zeros(len(rows), np.int64)
    left, right, var, val = tr["left"].astype(np.int64), tr["right"].astype(np.int64), tr["var"].astype(np.int64), tr["val"]
    lvl = 0
    handed = {c: None for c in caps}
    out = []
    while True:
        internal = left[node] != 0
        if not internal.any():
            break
        ids, cnt = np.unique(node[internal], return_counts=True)
        for c in caps:
            if handed[c] is None and cnt.max() <= c:
                handed[c] = (lvl, int(cnt.sum()), len(ids))
        out.append((lvl, len(ids), int(cnt.sum()), int(cnt.max()), int((cnt <= 32).sum()), int(cnt[cnt <= 32].sum()), int(cnt[cnt <= 128].sum())))
        go = X[rows, var[node]] <= val[node]
        node = np.where(internal, np.where(go, left[node], right[node]), node)
        lvl += 1
    if t < 2:
        print("tree %d: %d nodes, %d in-bag rows" % (t, len(left), len(rows)))
        for o in out:
            print("  level %2d: open %4d entries %5d max %5d | nodes<=32: %4d (entries %5d) | entries in nodes<=128: %5d" % o)
    print("tree %d hand-over (cap: level, entries, open nodes):" % t, {c: handed[c] for c in caps}, "levels", lvl, "entry-levels", sum(o[2] for o in out))
