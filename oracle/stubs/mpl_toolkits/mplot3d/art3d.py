class Poly3DCollection:  # placeholder type, never instantiated
    pass
