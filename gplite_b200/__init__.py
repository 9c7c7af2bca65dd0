"""gplite_b200 -- B200-native (sm_100a) implementation of gplite's dense GP hot path.

Host-side mirror of the R API (gp_init / gp_fit / gp_optim / gp_energy / gp_pred, cf_*, lik_*,
method_*, approx_*, prior_*) over libgplite_b200.so.  The CUDA library is the product; importing
this package never falls back to a CPU implementation.
"""
from .model import (cf_const, cf_lin, cf_sexp, cf_matern32, cf_matern52, cf_nn, cf_periodic, cf_prod,
                    lik_gaussian, lik_bernoulli, lik_binomial, lik_betabinom, lik_poisson,
                    prior_fixed, prior_logunif, prior_lognormal, prior_half_t,
                    method_full, method_fitc, method_rf, approx_laplace, approx_ep)
from .gp import (gp_init, gp_fit, gp_energy, gp_energy_grad, gp_optim, gp_pred, get_param,
                 get_param_vector, set_param, eval_cf, fit_member, lpdf_prior, is_fitted, GP,
                 gp_loo, gp_compare, gp_draw, gp_pred_cov, get_inducing, gp_save, gp_load)
from .quadrature import gauss_hermite_points
from ._lib import Context, default_context, GplError, NotPositiveDefinite, load as load_library

__all__ = [n for n in dir() if not n.startswith("_")]
