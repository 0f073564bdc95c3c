"""Shared test helpers: golden fixtures and oracle construction (test infrastructure)."""
from __future__ import annotations

import io
import json
import os

import numpy as np
import pandas as pd

from oracle import mope_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_tables(name="tables_small.npz"):
    z = np.load(os.path.join(GOLDEN, name))
    return {k: z[k] for k in z.files}


def load_cases():
    with open(os.path.join(GOLDEN, "cases.json")) as f:
        return json.load(f)


def load_outputs():
    z = np.load(os.path.join(GOLDEN, "outputs.npz"))
    return {k: z[k] for k in z.files}


def read_tsv(text):
    df = pd.read_csv(io.StringIO(text), sep="\t", comment="#", na_values=["NA"], header=0)
    return {c: df[c].values for c in df.columns}


def oracle_tables(tb):
    drift = orc.TransitionData(tb["drift"], tb["gens"], tb["us"], int(tb["N"]), tb["freqs"], tb["breaks"])
    bot = orc.TransitionDataBottleneck(tb["bot"], tb["nbs"], tb["bot_us"], int(tb["N"]))
    return drift, bot


def oracle_inference(case, tb, **extra):
    drift, bot = oracle_tables(tb)
    peds = [(read_tsv(p["data"]), p["tree"], read_tsv(p["ages"])) for p in case["peds"]]
    opts = dict(log_unif_drift=True)
    opts.update(case.get("opts", {}))
    opts.update(extra)
    return orc.Inference(peds, drift, bot, **opts)


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
