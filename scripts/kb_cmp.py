import sys, numpy as np
a, b = np.load(sys.argv[1]), np.load(sys.argv[2])
worst = 0
for n in a.files:
    d = np.linalg.norm(a[n] - b[n]); r = np.linalg.norm(b[n])
    rel = d / max(r, 1e-30)
    if ("conv1" in n or "conv2" in n or "bn1" in n or "bn2" in n or n == "logp") and not n.startswith("fc") or rel > 1e-3:
        print("%-28s rel %.3e  |b| %.3e" % (n, rel, r))
    worst = max(worst, rel if r > 1e-12 else 0)
print("worst rel", worst)
