#!/usr/bin/env python
"""Golden fixtures of the posterior-mutation-location likelihoods, produced by RUNNING THE REFERENCE (ammodramus/mope):

    python tests/golden/make_golden_pml.py        (build container only: needs /root/reference)

mope/posteriormutloc.py:48-271: `fill_cond_probs` (per-locus log-likelihoods of the pruning pass, and the same with the root
distribution restricted to the two fixed states, "mutation only") and `fill_cond_probs_mutation` (per-locus log-likelihood
that the heteroplasmy arose by mutation on the branch above a focal node: masked transition matrices below / at / beside
the focal node) on the fixtures of make_golden.py (tables_small.npz, the example data on mope_study_both.newick: age-scaled,
bottleneck and fixed-length branches).  Writes outputs_pml.npz."""
from __future__ import annotations

import json
import os
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle import install_ref  # noqa: E402

warnings.filterwarnings("ignore")


def main():
    install_ref.install()
    assert install_ref.activate()
    from make_golden_selection import write_master
    from mope import inference as refinf
    from mope import posteriormutloc as pml
    tb = {k: v for k, v in np.load(os.path.join(HERE, "tables_small.npz")).items()}
    with open(os.path.join(HERE, "cases.json")) as f:
        cases = json.load(f)
    out = {}
    with tempfile.TemporaryDirectory() as wd:
        dm = write_master(os.path.join(wd, "drift.h5"), tb["drift"], "gen", [int(g) for g in tb["gens"]], "u", [float(u) for u in tb["us"]],
                          int(tb["N"]), tb["freqs"], tb["breaks"])
        bm = write_master(os.path.join(wd, "bot.h5"), tb["bot"], "Nb", [int(g) for g in tb["nbs"]], "u", [float(u) for u in tb["bot_us"]],
                          int(tb["N"]), tb["freqs"], tb["breaks"])
        for name in ("example_both", "example_fixed"):
            case = cases[name]
            ped = case["peds"][0]
            fns = []
            for key in ("data", "tree", "ages"):
                fn = os.path.join(wd, "%s_%s.txt" % (name, key))
                with open(fn, "w") as f:
                    f.write(ped[key])
                fns.append(fn)
            I = refinf.Inference(data_files=[fns[0]], tree_files=[fns[1]], age_files=[fns[2]], transitions_file=dm, true_parameters=None,
                                 start_from="prior", data_are_freqs=False, bottleneck_file=bm, log_unif_drift=True, lower_drift_limit=1e-3,
                                 upper_drift_limit=3, genome_size=16569, poisson_like_penalty=1.0, min_freq=0.001, min_phred_score=None)
            rng = np.random.RandomState(23)
            X = rng.uniform(0.6 * I.lower + 0.4 * I.upper, 0.4 * I.lower + 0.6 * I.upper, (2, I.ndim))
            mrca = I.trees[0]
            nodes = [nd.name for nd in mrca.walk(mode="postorder") if nd != mrca]
            overall, mutonly, focal = [], [], []
            for x in X:
                I.transition_data.clear_cache()
                a, b = pml.fill_cond_probs(I, 0, x, include_mutation_only=True)
                overall.append(np.array(a)); mutonly.append(np.array(b))
                per = []
                for nd in mrca.walk(mode="postorder"):
                    if nd == mrca:
                        continue
                    I.transition_data.clear_cache()
                    per.append(np.array(pml.fill_cond_probs_mutation(nd, I, 0, x)))
                focal.append(per)
            out[name + "/X"] = X
            out[name + "/overall"] = np.array(overall)          # [n_x, L]
            out[name + "/mutation_only"] = np.array(mutonly)    # [n_x, L]
            out[name + "/focal"] = np.array(focal)              # [n_x, n_nodes, L]
            out[name + "/nodes"] = np.array(nodes)
            print(name, "nodes", nodes, "overall[0][:3]", overall[0][:3], "focal finite frac", np.isfinite(np.array(focal)).mean())
    np.savez_compressed(os.path.join(HERE, "outputs_pml.npz"), **out)
    print("wrote", os.path.join(HERE, "outputs_pml.npz"))


if __name__ == "__main__":
    main()
