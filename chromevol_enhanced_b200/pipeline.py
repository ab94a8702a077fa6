"""Host pipeline around the device likelihood path.

  quantify_rearrangements_from_tree   reference AR:1155-1217 (fusion / fission event table)
  write_ancestral_states_report       reference AR:1098-1128 (ancestral-states CSV)
  ancestral_reconstruction            reference AR:970-1152  (--use_chromevol branch only)
  main                                reference AR:1220-1327 (same flags)

Everything numeric runs through the drop-in classes of ``ancestral_reconstruction.py`` (CUDA);
this module is file I/O, naming rules and sorting.  Out of scope here, as in DESIGN.md section 8:
the parsimony fallback (AR:1038-1088), ``plot_annotated_tree`` (AR:815-926; needs ete3.treeview /
Qt) and ``analyze_rearrangements`` (AR:929-967; synteny TSV) -- requesting them logs a message and
the rest of the run continues, exactly as the reference continues when treeview is missing.
"""
from __future__ import annotations

import argparse
import logging
import os

import numpy as np
import pandas as pd

from .ancestral_reconstruction import (ChromEvolLikelihoodCalculator,
                                       calculate_model_selection_criteria,
                                       perform_chromevol_analysis)
from .newick import Tree

EVENT_COLUMNS = ["branch_to_node", "parent_node", "event_type", "change_magnitude", "count_from",
                 "count_to"]


def rearrangement_events(tree):
    """One row per branch whose two ends carry different ``count`` features (AR:1163-1185):
    a larger child count is a "fission", a smaller one a "fusion"."""
    rows = []
    for node in tree.traverse():
        if node.is_root():
            continue
        child = getattr(node, "count", None)
        parent = getattr(node.up, "count", None)
        if child is None or parent is None:
            continue
        delta = child - parent
        if delta == 0:
            continue
        rows.append({
            "branch_to_node": node.name if node.name else f"To_Leaf_{node.get_leaves()[0].name[:10]}",
            "parent_node": node.up.name if node.up.name else "Root_or_Unnamed",
            "event_type": "fission" if delta > 0 else "fusion",
            "change_magnitude": abs(delta),
            "count_from": parent,
            "count_to": child,
        })
    return rows


def quantify_rearrangements_from_tree(tree, output_file):
    """AR:1155-1217: event table sorted by (type ascending, magnitude descending), CSV, totals."""
    logging.info("\n--- Global Analysis of Chromosome Number Evolution (from tree structure) ---")
    rows = rearrangement_events(tree)
    if not rows:
        logging.info("No chromosome number changes detected across the tree.")
        if output_file:
            pd.DataFrame(columns=EVENT_COLUMNS).to_csv(output_file, index=False)
            logging.info(f"* Empty rearrangement events report saved to {output_file}")
        return None
    events = pd.DataFrame(rows)
    events.sort_values(by=["event_type", "change_magnitude"], ascending=[True, False], inplace=True)
    if output_file:
        events.to_csv(output_file, index=False)
        logging.info(f"* Tree-based rearrangement events report saved to {output_file}")
    logging.info("\nDetailed report of chromosomal changes per branch (from tree structure):")
    for _, ev in events.iterrows():
        logging.info(f" - Branch from {ev['parent_node']} to {ev['branch_to_node']}: {ev['event_type']} "
                     f"of {ev['change_magnitude']} (from {ev['count_from']} to {ev['count_to']})")
    fusions = events.loc[events["event_type"] == "fusion", "change_magnitude"].sum()
    fissions = events.loc[events["event_type"] == "fission", "change_magnitude"].sum()
    logging.info(f"\nTotal inferred fusion events (magnitude sum): {fusions}")
    logging.info(f"Total inferred fission events (magnitude sum): {fissions}")
    logging.info("\n--- Overall Evolutionary Trend (from tree structure) ---")
    if fissions > fusions:
        logging.info("The evolutionary history of this group appears to be dominated by chromosome fissions.")
    elif fusions > fissions:
        logging.info("The evolutionary history of this group appears to be dominated by chromosome fusions.")
    else:
        logging.info("Chromosome fusions and fissions appear to be relatively balanced across the tree.")
    return events


def ancestral_states_rows(tree, use_chromevol_method=True):
    """Rows of the ancestral-states report in the tree's default (level-order) traversal
    (AR:1098-1123).  Under --use_chromevol leaves are included and the columns are
    node_name, inferred_chromosome_count, ml_probability: the reference asks
    ``_summarize_event_counts({})`` for the event types, which is empty, so no expected_* column
    is ever written (SURVEY.md section 8(a) notes)."""
    rows = []
    for k, node in enumerate(tree.traverse()):
        name = node.name if node.name else f"Internal_Node_{k}"
        if node.is_leaf() and not node.name:
            name = f"Leaf_{node.get_leaves()[0].name}"
        row = {"node_name": name, "inferred_chromosome_count": getattr(node, "count", None)}
        if use_chromevol_method:
            row["ml_probability"] = getattr(node, "ml_probability", None)
        if not node.is_leaf() or use_chromevol_method:
            rows.append(row)
    return rows


def write_ancestral_states_report(tree, output_file, use_chromevol_method=True):
    frame = pd.DataFrame(ancestral_states_rows(tree, use_chromevol_method))
    frame.to_csv(output_file, index=False)
    logging.info(f"\n* Ancestral states report saved to {output_file}")
    return frame


def write_reconstructions_report(tree, counts_dict, model, root_prior_state, output_file):
    """ADDITIVE (no reference counterpart; AR:274-295 documents the missing down-pass): the marginal
    posterior arg max and the joint (Pupko) assignment of every node next to the reference's
    up-pass arg max, with the parameters the analysis ended with.  The reference-shaped report
    and the node attributes it reads are not touched."""
    calculator = ChromEvolLikelihoodCalculator(tree, counts_dict, model)
    root_freq = np.ones(model.max_chr + 1) / (model.max_chr + 1)
    if root_prior_state is not None and 0 <= root_prior_state <= model.max_chr:
        root_freq = np.zeros(model.max_chr + 1)
        root_freq[root_prior_state] = 1.0
    calculator.set_root_frequencies(root_freq)
    calculator.marginal_reconstruction()
    _, joint_logp = calculator.joint_reconstruction()
    rows = []
    for k, node in enumerate(tree.traverse()):
        name = node.name if node.name else f"Internal_Node_{k}"
        rows.append({"node_name": name,
                     "up_pass_count": getattr(node, "ml_count", None),
                     "marginal_count": int(node.marginal_count),
                     "marginal_probability": float(node.marginal_probability),
                     "joint_count": int(node.joint_count)})
    frame = pd.DataFrame(rows)
    frame.to_csv(output_file, index=False)
    logging.info(f"* Marginal / joint reconstructions saved to {output_file} "
                 f"(joint log-probability {joint_logp:.4f})")
    return frame


def ancestral_reconstruction(tree_file, counts_file, map_file, output_image, output_ancestral_states,
                             output_rearrangements, root_constraint_count, layout,
                             show_branch_length, show_support, analyze_pairs_list, dpi,
                             use_chromevol_method, optimize_model_params, perform_model_selection,
                             max_chromosome_val, initial_model_params_dict, output_reconstructions=None):
    """AR:970-1152 for the ChromEvol branch; returns the annotated tree (the reference returns
    None; callers that ignore the value see no difference)."""
    logging.info("--- Loading Data ---")
    logging.info(f"* Tree file: {tree_file}")
    logging.info(f"* Counts file: {counts_file}")
    if map_file:
        logging.info(f"* Map file: {map_file}")
    if not os.path.exists(tree_file):
        logging.error(f"Tree file not found: {tree_file}")
        return None
    if not os.path.exists(counts_file):
        logging.error(f"Counts file not found: {counts_file}")
        return None
    if map_file and not os.path.exists(map_file):
        logging.warning(f"Map file specified but not found: {map_file}")

    tree = Tree(tree_file, format=1)
    counts_df = pd.read_csv(counts_file)
    counts_dict = dict(zip(counts_df["species"], counts_df["chromosome_count"]))
    observed_max = counts_df["chromosome_count"].max()
    if max_chromosome_val is None:                                   # AR:995-998
        max_chromosome_val = observed_max + 10
        logging.info(f"Dynamically set max_chromosome_val to {max_chromosome_val} based on input data.")
    elif observed_max > max_chromosome_val:
        logging.warning(f"Max observed chromosome count ({observed_max}) is greater than specified "
                        f"max_chromosome_val ({max_chromosome_val}). Consider increasing it.")
    for leaf in tree:
        if leaf.name in counts_dict:
            leaf.add_features(count=counts_dict[leaf.name])
        else:
            logging.warning(f"No chromosome count data found for species: {leaf.name}. "
                            "Will be treated as missing data.")
            leaf.add_features(count=None)

    if not use_chromevol_method:
        logging.error("The parsimony fallback (AR:1038-1088) is outside the scope of this package; "
                      "run with --use_chromevol.")
        return None

    logging.info("\n--- Using ChromEvol-Inspired Maximum Likelihood Method ---")
    if perform_model_selection:
        # results are logged, not fed forward (AR:1018-1022)
        results = calculate_model_selection_criteria(tree, counts_dict, max_chromosome_val,
                                                     root_constraint_count)
        logging.info(f"Model selection results: {results}")
    model, tree = perform_chromevol_analysis(tree, counts_dict, max_chromosome_val,
                                         root_prior_state=root_constraint_count, use_ml=True,
                                         optimize_params=optimize_model_params,
                                         stochastic_mapping=True,
                                         model_params_input=initial_model_params_dict)

    logging.info("\nPhylogenetic tree with inferred ancestral chromosome numbers (text version):")
    try:
        logging.info(tree.get_ascii(attributes=["name", "count", "ml_probability"], show_internal=True))
    except Exception as exc:  # noqa: BLE001  (the reference swallows this too, AR:1094-1095)
        logging.warning(f"Could not print tree ASCII representation: {exc}")

    write_ancestral_states_report(tree, output_ancestral_states, use_chromevol_method=True)
    if output_reconstructions:
        write_reconstructions_report(tree, counts_dict, model, root_constraint_count, output_reconstructions)
    if analyze_pairs_list:
        logging.warning("--analyze_pairs (synteny map analysis, AR:929-967) is outside the scope of this "
                        "package; skipping.")
    quantify_rearrangements_from_tree(tree, output_rearrangements)
    logging.warning(f"Tree rendering (AR:815-926, ete3.treeview) is outside the scope of this package; "
                    f"{output_image} is not written.")
    return tree


def build_parser():
    """The reference's flags (AR:1228-1273), unchanged."""
    p = argparse.ArgumentParser(
        description="Ancestral chromosome reconstruction with ChromEvol-inspired maximum likelihood "
                    "on a B200 GPU.", formatter_class=argparse.RawTextHelpFormatter)
    p.add_argument('--tree', type=str, default='data/pruned_phylogeny.nwk',
                   help='Path to the phylogenetic tree file in Newick format.')
    p.add_argument('--counts', type=str, default='data/chromosome_counts.csv',
                   help='Path to the CSV file with species chromosome counts (species,chromosome_count).')
    p.add_argument('--map', type=str, default=None,
                   help='Path to the TSV file with chromosome mapping/synteny data (for --analyze_pairs).')
    p.add_argument('--out_image', type=str, default='results/annotated_phylogeny.svg',
                   help='Path of the phylogeny image (not rendered by this package).')
    p.add_argument('--out_ancestors', type=str, default='results/ancestral_states_report.csv',
                   help='Path to save the CSV report of inferred ancestral states.')
    p.add_argument('--out_rearrangements', type=str, default='results/rearrangement_events_report.csv',
                   help='Path to save the CSV report of rearrangement events (from tree structure).')
    p.add_argument('--root_count', type=int, default=None,
                   help='Constrain the ancestral chromosome number for the root of the tree (ML prior).')
    p.add_argument('--max_chr_number', type=int, default=None,
                   help='Maximum chromosome number to consider in the model. If None, inferred from data + buffer.')
    p.add_argument('--layout', type=str, default='circular', choices=['circular', 'rectangular'],
                   help='Tree layout style for visualization.')
    p.add_argument('--show_branch_length', action='store_true', default=False,
                   help='Show branch lengths in the tree visualization.')
    p.add_argument('--show_support', action='store_true', default=False,
                   help='Show branch support values in the tree visualization.')
    p.add_argument('--dpi', type=int, default=300, help='DPI for raster image formats.')
    p.add_argument('--analyze_pairs', nargs='*', metavar='SPECIES', default=[],
                   help='Pairs of species for synteny-based rearrangement analysis (requires --map).')
    p.add_argument('--use_chromevol', action='store_true', default=False,
                   help='Use ChromEvol-inspired maximum likelihood methods.')
    p.add_argument('--optimize_params', action='store_true', default=False,
                   help='Optimize model parameters using maximum likelihood (requires --use_chromevol).')
    p.add_argument('--model_selection', action='store_true', default=False,
                   help='Perform model selection using AIC/BIC (requires --use_chromevol).')
    p.add_argument('--out_reconstructions', type=str, default=None,
                   help='(additive, not in the reference) Also run the marginal (up-pass + down-pass) and joint '
                        '(Pupko) ancestral reconstructions with the final parameters and write them to this CSV.')
    return p


def main(argv=None):
    logging.basicConfig(level=logging.INFO, format='%(levelname)s: %(message)s')   # AR:16
    args = build_parser().parse_args(argv)
    for path in (args.out_image, args.out_ancestors, args.out_rearrangements):
        folder = os.path.dirname(path)
        if folder:
            os.makedirs(folder, exist_ok=True)
    if not os.path.exists(args.tree):
        logging.error(f"Input tree file not found: {args.tree}")
        return 1
    if not os.path.exists(args.counts):
        logging.error(f"Input counts file not found: {args.counts}")
        return 1
    if args.analyze_pairs and (not args.map or not os.path.exists(args.map)):
        logging.warning(f"Map file {args.map} is required for --analyze_pairs but not found or not "
                        "specified. Skipping pair analysis.")
        args.analyze_pairs = []
    tree = ancestral_reconstruction(
        tree_file=args.tree, counts_file=args.counts, map_file=args.map, output_image=args.out_image,
        output_ancestral_states=args.out_ancestors, output_rearrangements=args.out_rearrangements,
        root_constraint_count=args.root_count, layout=args.layout,
        show_branch_length=args.show_branch_length, show_support=args.show_support,
        analyze_pairs_list=args.analyze_pairs, dpi=args.dpi, use_chromevol_method=args.use_chromevol,
        optimize_model_params=args.optimize_params, perform_model_selection=args.model_selection,
        max_chromosome_val=args.max_chr_number, initial_model_params_dict=None,
        output_reconstructions=args.out_reconstructions)
    return 0 if tree is not None else 1


if __name__ == "__main__":
    raise SystemExit(main())
