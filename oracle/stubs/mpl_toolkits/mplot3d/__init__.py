class Axes3D:  # placeholder type, never instantiated
    pass
