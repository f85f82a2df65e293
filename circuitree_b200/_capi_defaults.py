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


def i_step(n_reactions: float, n_species: float) -> float:
    """Algorithmic issue slots (lane-instructions) of one SSA step of the pairwise model with
    ``n_reactions`` reactions and ``n_species`` genes (SPEC.md section 8): half a Philox4x32-10 block
    (30) + per reaction the propensity fused into the running sum, one compare and 0.75 state
    updates (2.75 R) + rate selects (7 M) + tau, grid check and loop (13)."""
    return 30.0 + 2.75 * n_reactions + 7.0 * n_species + 13.0


def i_step_dimers(n_entries: float, n_tf: float) -> float:
    """Same for the dimer model (SPEC.md sections 8-9) with ``n_entries`` reactions and ``n_tf`` TFs:
    half a Philox block (30) + per reaction one operand read, the propensity fused into the running
    sum and its share of the search (3 R) + per TF the transcription-rate and (un)binding-rate
    selects (7 n_tf) + tau, grid check, loop (13) + firing: four state reads, four adds, four
    writes (12)."""
    return 30.0 + 3.0 * n_entries + 7.0 * n_tf + 13.0 + 12.0
