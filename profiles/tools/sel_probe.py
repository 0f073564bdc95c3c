"""Selection-model throughput probe (parity path, one parameter vector per call): the example data replicated to L loci
with P distinct positions, the six-branch tree of tests/golden/cases_sel.json, the F = 65 selection tables of the fixtures."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from tests import helpers as H
from mope_b200 import Inference, TransitionTables

if os.environ.get("SEL_PROBE_F201"):
    # the F = 201 fixture's drift matrices with their rate axis relabelled as selection coefficients (a valid table set for timing)
    TB = dict(H.load_tables("tables_f201.npz"))
    TB["us"] = np.array([-6e-3, 6e-3])
else:
    TB = H.load_tables("tables_sel.npz")
case = json.load(open(os.path.join(H.GOLDEN, "cases_sel.json")))["sel_example"]
ped = case["peds"][0]
lines = ped["data"].splitlines()
hdr, rows = lines[0], lines[1:]
out = {}
for reps, npos in ((1, 40), (20, 400), (100, 2000)):
    rng = np.random.RandomState(3)
    body = []
    for r in range(reps):
        for l in rows:
            parts = l.split("\t")
            parts[-1] = str(73 + 16 * rng.randint(npos))
            body.append("\t".join(parts))
    data = "\n".join([hdr] + body) + "\n"
    tables = TransitionTables(drift=TB["drift"], gens=TB["gens"], us=TB["us"], N=float(TB["N"]), freqs=TB["freqs"], breaks=TB["breaks"])
    inf = Inference([data], [ped["tree"]], [ped["ages"]], tables, log_unif_drift=True, selection_model=True, _inputs_are_text=True)
    rng = np.random.RandomState(5)
    x = rng.uniform(0.7 * inf.lower + 0.3 * inf.upper, 0.3 * inf.lower + 0.7 * inf.upper)
    P = inf.num_unique_positions
    x[inf.num_varnames:2 * inf.num_varnames] = 0.0
    x[-P:] = rng.choice([-1.0, 1.0], P) * (10.0 + rng.uniform(0, 4, P)) * (rng.rand(P) < 0.7)
    for _ in range(2):
        ll = inf.loglike(x.copy())
    torch.cuda.synchronize(); t0 = time.perf_counter(); n = 5
    for _ in range(n):
        ll = inf.loglike(x.copy())
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
    L = inf.num_loci[0]
    out["L%d_P%d" % (L, P)] = dict(loci=L, positions=P, branches=inf.num_branches[0], F=int(tables.F), ms_per_eval=dt * 1e3,
                                   loci_per_s=L / dt, ll=float(ll))
    print(out["L%d_P%d" % (L, P)], flush=True)
    inf.close()
json.dump(out, open(os.path.join("gpurun_out", "r02_sel_probe%s%s.json" % ("_f201" if os.environ.get("SEL_PROBE_F201") else "",
                                                                             "_unfused" if os.environ.get("MOPE_SEL_UNFUSED") else "")), "w"), indent=1)
