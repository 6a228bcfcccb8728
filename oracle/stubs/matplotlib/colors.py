class LinearSegmentedColormap:
    @staticmethod
    def from_list(name, colors, *a, **k):
        return (name, tuple(colors))
