"""tile vs plain residual norm against a long-double sum of the stored residual"""
import json, math, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multigrid_parallel_b200 as m
gold = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/histories.json")))
for levels in (9,):
    g = gold[f"3_{levels}_2"]["history"]
    for graph in (1, 0):
        s = m.Solver(3, levels, 2)
        s.set_option(0, graph)
        top = levels - 1
        s.set_dirichlet(top, m.MGB_D)
        s.set_dirichlet(top, m.MGB_U)
        for cyc in range(16):
            n_cyc = s.vcycle()
            n_tile = s.residual(top, store_r=False)
            m.set_global(m.G_TILE, 0)
            n_plain = s.residual(top, store_r=False)
            m.set_global(m.G_TILE, 1)
            print(f"graph={graph} cyc={cyc+1} cycle-norm dev from plain {(n_cyc-n_plain)/n_plain:+.2e} tile {(n_tile-n_plain)/n_plain:+.2e} golden {(g[cyc]-n_plain)/n_plain:+.2e}", flush=True)
        s.close()
