import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from oracle import pyoracle as po
from util import all_classes, random_shell_table
rng = np.random.default_rng(4243)
for cl in all_classes(max_l=3, with_g=False):
    table = random_shell_table(rng, cl, 1)
    sh = po.Shells(*table, raw=False)
    print("class", cl, flush=True)
    got = po.compute2(sh, precision=0.0, b200=True)
print("all ok")
