"""SolverConfig -- the reference's configuration object, kept field for field
(pyhype/solver_config.py:32-138)."""
from __future__ import annotations

from .states import ConservativeState


class SolverConfig:
    __slots__ = [
        "fvm_type", "fvm_spatial_order", "fvm_num_quadrature_points", "fvm_gradient_type",
        "fvm_flux_function_type", "fvm_slope_limiter_type", "time_integrator", "initial_condition",
        "interface_interpolation", "reconstruction_type", "write_solution", "write_solution_mode",
        "write_solution_name", "write_solution_base", "write_every_n_timesteps", "plot_every",
        "plot_function", "CFL", "t_final", "realplot", "profile", "fluid", "nx", "ny", "n", "nghost",
        "use_JIT", "show_log_for_procs",
    ]

    def __init__(
        self, nx, ny, CFL, t_final, initial_condition, fvm_type, time_integrator, fvm_gradient_type,
        fvm_flux_function_type, fvm_slope_limiter_type, fvm_spatial_order, fvm_num_quadrature_points, fluid,
        nghost=1, use_JIT=True, profile=False, realplot=False, plot_every=20, plot_function="Density",
        write_solution=False, write_solution_mode="every_n_timesteps", write_solution_name="nozzle",
        write_solution_base=None, reconstruction_type=ConservativeState, write_every_n_timesteps=40,
        interface_interpolation="arithmetic_average", show_log_for_procs=None,
    ):
        self.initial_condition = initial_condition
        self.nx, self.ny, self.n, self.nghost = nx, ny, nx * ny, nghost
        self.CFL, self.t_final = CFL, t_final
        self.fvm_type = fvm_type
        self.time_integrator = time_integrator
        self.fvm_gradient_type = fvm_gradient_type
        self.fvm_flux_function_type = fvm_flux_function_type
        self.fvm_slope_limiter_type = fvm_slope_limiter_type
        self.fvm_spatial_order = fvm_spatial_order
        self.fvm_num_quadrature_points = fvm_num_quadrature_points
        self.reconstruction_type = reconstruction_type
        self.interface_interpolation = interface_interpolation
        self.fluid = fluid
        self.use_JIT, self.profile, self.realplot = use_JIT, profile, realplot
        self.plot_every, self.plot_function = plot_every, plot_function
        self.write_solution = write_solution
        self.write_solution_mode = write_solution_mode
        self.write_solution_name = write_solution_name
        self.write_solution_base = write_solution_base
        self.write_every_n_timesteps = write_every_n_timesteps
        if show_log_for_procs is None:
            self.show_log_for_procs = [0]
        elif show_log_for_procs == "all":
            self.show_log_for_procs = "all"
        else:
            self.show_log_for_procs = show_log_for_procs

    def __str__(self):
        return "".join(f"\t{atr}: {getattr(self, atr)}\n" for atr in self.__slots__)
