"""Import-only stand-in for emcee (test infrastructure; the sampler is never run by the oracle)."""


class EnsembleSampler(object):
    def __init__(self, *a, **k):
        raise RuntimeError("emcee is not installed; this is an import-only stub")


class PTSampler(EnsembleSampler):
    pass
