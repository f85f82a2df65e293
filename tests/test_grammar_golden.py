"""Host grammars vs outputs of the unmodified reference (tests/golden/make_golden.py).

Covers SURVEY.md section 8 rows a2, a3, a8, a9: state strings, action order, canonical forms,
has_pattern and parse_genotype are bit-exact with reference circuitree/models.py.
"""
import numpy as np
import pytest

from circuitree_b200 import DimersGrammar, SimpleNetworkGrammar

AI = ["activates", "inhibits"]


def grammar_for(state, cache={}):
    n = len(state.strip("*").split("::")[0])
    if n not in cache:
        cache[n] = SimpleNetworkGrammar(list("ABCDE"[:n]), AI)
    return cache[n]


def test_actions_order_matches_reference(ref_vectors):
    g3 = SimpleNetworkGrammar(list("ABC"), AI)
    g5 = SimpleNetworkGrammar(list("ABCDE"), AI)
    for state, actions in ref_vectors["actions3"].items():
        assert g3.get_actions(state) == actions, state
    for state, actions in ref_vectors["actions5"].items():
        assert g5.get_actions(state) == actions, state


def test_root_without_interactions_offers_no_terminate(ref_vectors):
    # quirk of reference models.py:198-204 that must be preserved
    g3 = SimpleNetworkGrammar(list("ABC"), AI)
    acts = g3.get_actions("ABC::")
    assert acts == ref_vectors["actions3"]["ABC::"]
    assert "*terminate*" not in acts and len(acts) == 18


def test_actions_max_interactions(ref_vectors):
    g = SimpleNetworkGrammar(list("ABC"), AI, max_interactions=4)
    for state, actions in ref_vectors["actions3_max4"].items():
        assert g.get_actions(state) == actions, state


def test_actions_partial_components_as_sets(ref_vectors):
    # the reference lists component additions in set order (hash-seed dependent)
    g = SimpleNetworkGrammar(list("ABC"), AI, root="A::")
    for state, actions in ref_vectors["actions_partial"].items():
        mine = g.get_actions(state)
        assert sorted(mine) == sorted(actions), state
        edges = [a for a in actions if len(a) == 3]
        assert [a for a in mine if len(a) == 3] == edges, state


def test_unique_state_matches_reference(ref_vectors):
    for state, unique in ref_vectors["unique"].items():
        assert grammar_for(state).get_unique_state(state) == unique, state
    for state, unique in ref_vectors["unique_partial"].items():
        assert SimpleNetworkGrammar(list("ABC"), AI).get_unique_state(state) == unique, state


def test_unique_state_fixed_components(ref_vectors):
    g = SimpleNetworkGrammar(list("ABCD"), AI, fixed_components=["A"])
    for state, unique in ref_vectors["unique_fixedA"].items():
        assert g.get_unique_state(state) == unique, state


def test_unique_state_is_idempotent_and_cache_independent(ref_vectors):
    small = SimpleNetworkGrammar(list("ABCD"), AI, cache_maxsize=2)
    nocache = SimpleNetworkGrammar(list("ABCD"), AI, cache_maxsize=None)
    for state, unique in ref_vectors["unique"].items():
        if len(state.strip("*").split("::")[0]) != 4:
            continue
        assert small.get_unique_state(state) == unique
        assert nocache.get_unique_state(unique) == unique


def test_do_action_docstring_examples(ref_vectors):
    g = SimpleNetworkGrammar(list("ABC"), AI)
    for state, action, expected in ref_vectors["do_action"]:
        assert g.do_action(state, action) == expected
    with pytest.raises(ValueError):
        g.do_action("ABC::", "AB")


def test_has_pattern(ref_vectors):
    g = SimpleNetworkGrammar(list("ABC"), AI)
    pats = ref_vectors["has_pattern"]["patterns"]
    for state, flags in ref_vectors["has_pattern"]["cases"]:
        assert [g.has_pattern(state, p) for p in pats] == flags, state
    with pytest.raises(ValueError):
        g.has_pattern("ABC::ABa", "ABC::ABa")
    with pytest.raises(ValueError):
        g.has_pattern("ABa", "ABa")


def test_parse_genotype(ref_vectors):
    for state, (comps, act, inh) in ref_vectors["parse_genotype"].items():
        c, a, i = SimpleNetworkGrammar.parse_genotype(state, nonterminal_ok=True)
        assert c == comps
        assert a.dtype == np.int_ and i.dtype == np.int_
        assert a.tolist() == act and i.tolist() == inh
    with pytest.raises(ValueError):
        SimpleNetworkGrammar.parse_genotype("ABC::ABa")


def test_to_dict_matches_reference(ref_vectors):
    g = SimpleNetworkGrammar(components=["A", "B", "C"], interactions=AI, root="ABC::")
    assert g.to_dict() == ref_vectors["to_dict_simple"]
    g.get_actions("ABC::ABa")
    g.get_unique_state("ABC::BAa")
    assert g.to_dict() == ref_vectors["to_dict_simple"]
    d = DimersGrammar(components=["A", "B"], regulators=["R"], interactions=AI)
    assert d.to_dict() == ref_vectors["dimers"]["to_dict"]


def test_constructor_validation():
    with pytest.raises(ValueError):
        SimpleNetworkGrammar(["Ab", "Ac"], AI)
    with pytest.raises(ValueError):
        SimpleNetworkGrammar(["A", "B"], ["activates", "attenuates"])


def test_pickle_roundtrip():
    import pickle

    g = SimpleNetworkGrammar(list("ABC"), AI)
    g.get_unique_state("ABC::BAa")
    g2 = pickle.loads(pickle.dumps(g))
    assert g2.get_unique_state("*CBA::CBi_ACa") == g.get_unique_state("*CBA::CBi_ACa")


def test_undo_actions_roundtrip():
    g = SimpleNetworkGrammar(list("ABC"), AI, root="ABC::")
    s = "ABC::ABa_BCi"
    undo = g.get_undo_actions(s)
    assert set(undo) == {"ABa", "BCi"}
    assert g.undo_action(s, "BCi") == "ABC::ABa"
    assert g.get_undo_actions("*" + s) == ["*undo_terminate*"]
    assert g.undo_action("*" + s, "*undo_terminate*") == s
    assert g.get_undo_actions("ABC::") == []


# ---- DimersGrammar: only the parts that run upstream are pinned ----------
@pytest.fixture()
def dimers():
    return DimersGrammar(components=["A", "B"], regulators=["R"], interactions=AI)


def test_dimers_do_action_and_parts(ref_vectors, dimers):
    for state, action, expected in ref_vectors["dimers"]["do_action"]:
        assert dimers.do_action(state, action) == expected
    for state, parts in ref_vectors["dimers"]["genotype_parts"]:
        assert list(dimers.get_genotype_parts(state)) == parts
    with pytest.raises(ValueError):
        dimers.do_action("AB+R::", "ABC")


def test_dimers_parse_genotype(ref_vectors, dimers):
    for state, (c, r, act, inh) in ref_vectors["dimers"]["parse_genotype"].items():
        cc, rr, a, i = dimers.parse_genotype(state)
        assert (cc, rr) == (c, r)
        assert a.tolist() == act and i.tolist() == inh


def test_dimers_options_and_actions(ref_vectors, dimers):
    assert dimers.dimer_options == ref_vectors["dimers"]["dimer_options"]
    assert sorted(sorted(g) for g in dimers.edge_options) == ref_vectors["dimers"]["edge_options_sorted"]
    for state, actions in ref_vectors["dimers"]["actions_sorted"].items():
        assert sorted(dimers.get_actions(state)) == actions, state
    d2 = DimersGrammar(components=["A", "B"], regulators=["R"], interactions=AI, max_interactions=2)
    for state, actions in ref_vectors["dimers"]["actions_sorted_max2"].items():
        assert sorted(d2.get_actions(state)) == actions, state


def test_dimers_unique_state_documented_intent(dimers):
    # restatement of reference models.py:759-778 (the upstream code raises on every input)
    a = dimers.get_unique_state("*BA+R::RBiA_AAaA")
    b = dimers.get_unique_state("*AB+R::BBaB_ARiB")
    assert a == b == "*AB+R::AAaA_BRiA"
    assert dimers.get_unique_state(a) == a
    assert dimers.get_unique_state("AB+R::") == "AB+R::"
    # zero-interaction states offer interactions only, like SimpleNetworkGrammar
    acts = dimers.get_actions("AB+R::")
    assert acts and all(len(x) == 4 for x in acts)
    assert dimers.has_pattern("*AB+R::ARiB_BBaB", "AAaA")
    # like the reference (models.py:927-937) the test is on (dimer, target) triplets, not on the type
    assert dimers.has_pattern("*AB+R::ARiB_BBaB", "AAiA")
    assert not dimers.has_pattern("*AB+R::ARiB_BBaB", "AAaA_BBaB")


def test_dimers_search_space_is_finite_and_canonical(dimers):
    from circuitree_b200 import CircuiTree

    class T(CircuiTree):
        def get_reward(self, state, **kw):
            return 0.0

    g = DimersGrammar(components=["A", "B"], regulators=["R"], interactions=AI, max_interactions=2)
    t = T(grammar=g, root="AB+R::", seed=0)
    t.grow_tree()
    assert all(g.get_unique_state(n) == n for n in t.graph.nodes)
    assert any(n.startswith("*") for n in t.graph.nodes)
