class LineCollection:
    def __init__(self, *a, **k):
        pass
