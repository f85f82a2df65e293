"""SPEC.md defaults as plain Python (kept equal to ``ct_default_config`` by tests/test_capi.py)."""

SPEC_DEFAULTS = {
    "nt": 2000,
    "decim": 16,
    "dt": 20.0,
    "init_mean": 10.0,
    # (lo, hi, is_log) for k_on, k_off1, k_off2, km_unbound, km_act, km_rep, km_act_rep, k_p, gamma_m, gamma_p
    "ranges": [
        (-2.3, -1.7, 1.0), (-1.0, -0.4, 1.0), (-2.7, -2.1, 1.0), (-0.3, 0.3, 1.0), (0.0, 0.6, 1.0),
        (-3.5, -2.5, 1.0), (-1.3, -0.7, 1.0), (-1.9, -1.3, 1.0), (-2.0, -1.4, 1.0), (-2.7, -2.1, 1.0),
    ],
    # dimer model only (SPEC.md section 9): k_dim, k_undim
    "dimer_ranges": [(-5.3, -4.7, 1.0), (-2.8, -2.2, 1.0)],
}
