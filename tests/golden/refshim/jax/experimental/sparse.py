def random_bcoo(*a, **k):
    raise NotImplementedError("sparse SLiCE variant is outside the hot path")
