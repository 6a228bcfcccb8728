"""Butcher tableaux of the explicit Runge-Kutta integrators the reference ships
(pyhype/time_marching/explicit_runge_kutta.py:91-274, factory at time_marching/factory.py:29-62).
Rows are lower-triangular ``a[s][k]``; the last row plays the role of the ``b`` weights.  The
coefficients are written with the reference's own Python expressions so the doubles agree."""

# Not in the reference's factory: the Heun / SSP-RK2 tableau that its README (and BASELINE.json's DMR
# config) name.  Offered under its own key so that "RK2" keeps the reference's midpoint meaning.
EXTRA_TABLEAUX = {
    "SSPRK2": [[1], [0.5, 0.5]],
}

TABLEAUX = {
    "ExplicitEuler1": [[1]],
    "RK2": [[0.5], [0, 1]],  # midpoint rule (the factory has no Heun / SSP-RK2)
    "Ralston2": [[2 / 3], [1 / 4, 3 / 4]],
    "RK3": [[0.5], [-1, 2], [1 / 6, 2 / 3, 1 / 6]],
    "RK3SSP": [[1], [1 / 4, 1 / 4], [1 / 6, 1 / 6, 2 / 3]],
    "Ralston3": [[1 / 2], [0, 3 / 4], [2 / 9, 1 / 3, 4 / 9]],
    "RK4": [[0.5], [0, 0.5], [0, 0, 1], [1 / 6, 1 / 3, 1 / 3, 1 / 6]],
    "Ralston4": [
        [0.4],
        [0.29697761, 0.15875964],
        [0.21810040, -3.05096516, 3.83286476],
        [0.17476028, -0.55148066, 1.20553560, 0.17118478],
    ],
    "DormandPrince5": [
        [1 / 5],
        [3 / 40, 9 / 40],
        [44 / 45, -56 / 15, 32 / 9],
        [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
        [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
        [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
    ],
}


def get_tableau(name: str):
    if name in ("Generic2", "Generic3"):
        # pyhype/time_marching/explicit_runge_kutta.py:114,155 read config.alpha, which is not a
        # SolverConfig slot: the reference raises AttributeError for these two names.
        raise AttributeError("'SolverConfig' object has no attribute 'alpha'")
    if name in EXTRA_TABLEAUX:
        return EXTRA_TABLEAUX[name]
    if name not in TABLEAUX:
        raise ValueError(f"Time marching scheme {name} is not available.")
    return TABLEAUX[name]
