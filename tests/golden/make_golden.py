"""Generate tests/golden/*.npz|json by running the UNMODIFIED reference module
(/root/reference/src/ancestral_reconstruction.py) in the build container, under the ete3
stand-in in tests/golden/_ete3_stub.  The reference has no tests of its own,
so these files -- plus the live comparison in tests/test_oracle_vs_reference.py -- are what
pins the oracle.  Re-run with:  python tests/golden/make_golden.py

Every file records the numpy / scipy versions that produced it (scipy is unpinned upstream).
"""
import io
import json
import os
import sys

import numpy as np
import pandas as pd
import scipy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from refload import REF_ROOT, load_reference  # noqa: E402

AR = load_reference()
from ete3 import Tree  # noqa: E402  (the stand-in)

from chromevol_enhanced_b200 import synthetic  # noqa: E402

VERSIONS = {"numpy": np.__version__, "scipy": scipy.__version__,
            "reference": "Biols9527/ChromEvol-Enhanced src/ancestral_reconstruction.py"}

OPT_FULL = dict(gain=2.1342596548720283, loss=0.2677580853199203, dupl=2.016554428743558,
                demidupl_gain=0.38097654656835955, demidupl_loss=1.1888333123910255,
                base_number_gain=0.01, base_number_loss=0.01, linear_rate=-0.03935220722444711)
OPT_GL = dict(gain=1.9823905216545576, loss=4.424978107995602, dupl=0, demidupl_gain=0,
              demidupl_loss=0, base_number_gain=0, base_number_loss=0, linear_rate=0)


def shipped_inputs():
    newick = open(os.path.join(REF_ROOT, "data", "pruned_phylogeny.nwk")).read().strip()
    df = pd.read_csv(os.path.join(REF_ROOT, "data", "chromosome_counts.csv"))
    counts = {str(k): int(v) for k, v in zip(df["species"], df["chromosome_count"])}
    return newick, counts


def run_case(newick, counts, max_chr, custom_params, root_state=None, want_nodes=True):
    tree = Tree(newick, format=1)
    model = AR.ChromEvolutionModel(max_chr, custom_params=custom_params)
    calc = AR.ChromEvolLikelihoodCalculator(tree, counts, model)
    n = max_chr + 1
    if root_state is None:
        freq = np.ones(n) / n
    else:
        freq = np.zeros(n)
        freq[root_state] = 1.0
    calc.set_root_frequencies(freq)
    ll = calc.calculate_total_log_likelihood()
    calc.ancestral_state_reconstruction()
    nodes = list(tree.traverse())
    out = {"logL": float(ll)}
    if want_nodes:
        out["vectors"] = np.stack([calc.likelihoods[id(nd)] for nd in nodes])
        out["ml_count"] = np.array([int(nd.ml_count) for nd in nodes])
        out["ml_probability"] = np.array([float(nd.ml_probability) for nd in nodes])
    return out, tree


def main():
    newick, counts = shipped_inputs()
    blobs = {}
    meta = {"versions": VERSIONS, "shipped_newick": newick, "shipped_counts": counts, "kat": {}}

    # --- Q matrices (N = 35) and structure checks (KAT 7)
    for tag, kw in (("default", {}), ("bn7", {"base_number": 7}), ("opt_full", OPT_FULL),
                    ("gl", OPT_GL), ("neg_lr", {"linear_rate": -0.5}), ("bn1", {"base_number": 1}),
                    ("bn2", {"base_number": 2})):
        blobs[f"Q35_{tag}"] = AR.ChromEvolutionModel(35, custom_params=kw).build_transition_rate_matrix()
    for N in (100, 256):
        Q = AR.ChromEvolutionModel(N, custom_params={"base_number": 5}).build_transition_rate_matrix()
        blobs[f"Q{N}_bn5_diag"] = np.diag(Q).copy()
        blobs[f"Q{N}_bn5_rowsum_abs"] = np.abs(Q).sum(axis=1)

    # --- P matrices
    for tag, kw in (("default", {}), ("opt_full", OPT_FULL), ("gl", OPT_GL)):
        m = AR.ChromEvolutionModel(35, custom_params=kw)
        for t in (0.0, 1e-6, 0.00299, 0.05, 0.3, 0.758285):
            blobs[f"P35_{tag}_t{t}"] = m.get_transition_probabilities(t)
    m100 = AR.ChromEvolutionModel(100, custom_params=OPT_FULL)
    for t in (0.01, 0.2):
        P = m100.get_transition_probabilities(t)
        blobs[f"P100_opt_full_t{t}_rows"] = P[[0, 1, 20, 50, 100]]

    # --- KATs on the shipped data
    kats = {
        "kat1_default_N35": (35, None, None),
        "kat2_N60_root24": (60, None, 24),
        "kat3_opt_full_N35": (35, OPT_FULL, None),
        "kat4_gainloss_N35": (35, OPT_GL, None),
        "kat5_bn7_N35": (35, {"base_number": 7}, None),
    }
    for name, (N, params, root_state) in kats.items():
        out, tree = run_case(newick, counts, N, params, root_state)
        meta["kat"][name] = {"max_chr": N, "params": params, "root_state": root_state,
                             "logL": out["logL"]}
        blobs[f"{name}_vectors"] = out["vectors"]
        blobs[f"{name}_ml_count"] = out["ml_count"]
        blobs[f"{name}_ml_probability"] = out["ml_probability"]
        if name == "kat3_opt_full_N35":
            for nd in tree.traverse():
                nd.add_features(count=nd.ml_count)
            buf = io.StringIO()
            AR.quantify_rearrangements_from_tree(tree, buf)
            meta["kat3_rearrangements_csv"] = buf.getvalue()
            meta["kat3_node_names"] = [nd.name for nd in tree.traverse()]
    meta["kat"]["kat3_aic_bic"] = {"AIC": 2 * 8 - 2 * meta["kat"]["kat3_opt_full_N35"]["logL"],
                                   "BIC": float(np.log(38) * 8 - 2 * meta["kat"]["kat3_opt_full_N35"]["logL"])}

    # --- KAT 6: the four-taxon example of docs/user_guide.md:37-46
    nw4 = "((A:0.1,B:0.1):0.05,(C:0.1,D:0.1):0.05);"
    c4 = {"A": 22, "B": 24, "C": 20, "D": 18}
    out, _ = run_case(nw4, c4, 34, None)
    meta["kat"]["kat6_four_taxa"] = {"newick": nw4, "counts": c4, "max_chr": 34, "logL": out["logL"]}
    blobs["kat6_vectors"] = out["vectors"]
    blobs["kat6_ml_count"] = out["ml_count"]

    # --- edge cases: polytomy + missing species + out-of-range count + zero / missing lengths
    nw_edge = "((A:0.1,B:0.2,C:0.05)X:0.3,(D:0,E)Y:0.01,F:0.4)R;"
    c_edge = {"A": 5, "B": 7, "C": 99, "E": 6, "F": 4}
    out, _ = run_case(nw_edge, c_edge, 12, {"base_number": 3})
    meta["kat"]["edge_polytomy"] = {"newick": nw_edge, "counts": c_edge, "max_chr": 12,
                                    "params": {"base_number": 3}, "logL": out["logL"]}
    blobs["edge_vectors"] = out["vectors"]
    blobs["edge_ml_count"] = out["ml_count"]

    # --- synthetic 200-taxon tree, N = 60 (deep enough to exercise many levels)
    tree = synthetic.random_tree(200, seed=11)
    cnt = synthetic.simulate_tip_counts(tree, 60, seed=12)
    nw = tree.write()
    for tag, params in (("default", None), ("opt_full", OPT_FULL)):
        out, _ = run_case(nw, cnt, 60, params)
        meta["kat"][f"synth200_{tag}"] = {"newick": nw, "counts": cnt, "max_chr": 60,
                                          "params": params, "logL": out["logL"]}
        blobs[f"synth200_{tag}_ml_count"] = out["ml_count"]
        blobs[f"synth200_{tag}_root_vector"] = out["vectors"][0]

    # --- stochastic mapping: reference replicate statistics on the shipped tree at the optimum
    np.random.seed(20261019)
    _, tree = run_case(newick, counts, 35, OPT_FULL)
    model = AR.ChromEvolutionModel(35, custom_params=OPT_FULL)
    R = 4000
    mapper = AR.StochasticMapper(tree, model, num_mappings=R)
    summary = mapper.perform_stochastic_mapping()
    nodes = list(tree.traverse())
    mean = np.zeros((len(nodes), 8))
    std = np.zeros((len(nodes), 8))
    tot = np.zeros((len(nodes), 8))
    types = ['gain', 'loss', 'duplication', 'base_number_gain', 'base_number_loss',
             'demi_gain_like', 'demi_loss_like', 'other']
    for k, nd in enumerate(nodes):
        if nd.is_root():
            continue
        for e, et in enumerate(types):
            mean[k, e] = summary[id(nd)][et]['mean']
            std[k, e] = summary[id(nd)][et]['std']
            tot[k, e] = summary[id(nd)][et]['total_occurrences']
    blobs["map_mean"], blobs["map_std"], blobs["map_total"] = mean, std, tot
    blobs["map_root_probs"] = np.asarray(tree.state_probabilities)
    meta["mapping"] = {"replicates": R, "max_chr": 35, "params": OPT_FULL, "numpy_seed": 20261019}

    np.savez_compressed(os.path.join(HERE, "golden.npz"), **blobs)
    with open(os.path.join(HERE, "golden.json"), "w") as fh:
        json.dump(meta, fh, indent=1)
    print("wrote", len(blobs), "arrays;", os.path.getsize(os.path.join(HERE, "golden.npz")), "bytes")
    for k, v in meta["kat"].items():
        print(k, v.get("logL", v))


if __name__ == "__main__":
    main()
