"""Global coupled kinase-TF-protein network path."""
from .network import GlobalSystem, synthetic_system, synthetic_loss_data  # noqa: F401
from .simulate import (LOSS_FN, fold_change_tables, metric_time_indices, simulate_and_measure, simulate_batch,  # noqa: F401
                       simulate_odeint, solve_custom, solve_custom_batch)
from .optproblem import GlobalODE_MOO, init_raw_params, unpack_params  # noqa: F401
from .sensitivity import compute_bounds, run_sensitivity_analysis  # noqa: F401
from .analysis import final_rate_of_change, simulate_until_steady, steady_check_batch, steady_time_grid  # noqa: F401
from .model_ivp import (make_solve_ivp_fun, make_solve_ivp_fun_combinatorial, make_solve_ivp_fun_distributive,  # noqa: F401
                        make_solve_ivp_fun_saturating, make_solve_ivp_fun_sequential)
