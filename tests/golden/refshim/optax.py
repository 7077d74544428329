def __getattr__(name):  # imported by train.py; the optimiser is not part of the golden vectors
    def _unused(*a, **k):
        raise NotImplementedError("optax is not shimmed")
    return _unused
