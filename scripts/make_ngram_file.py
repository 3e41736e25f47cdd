import numpy as np, sys
rng = np.random.default_rng(0)
n, N, V, D = 4, 8192 + 300, 500, 120
words = rng.integers(0, V, size=(N, n)); docs = (words[:, 0] * 3 + words[:, 1]) % D
with open(sys.argv[1], "w") as f:
    for i in range(N):
        f.write("%d %s%s\n" % (docs[i], " ".join(map(str, words[i])), " | 1.5" if i % 7 == 0 else ""))
