#!/usr/bin/env python
"""Write small sample files in the layouts the reference ships under ``examples/data`` (Kaggle MNIST CSV:
``label,pixel0..pixel783``; ATLAS Higgs CSV: ``EventId`` + 30 physics features + ``Weight`` + ``Label`` with
values ``s`` / ``b``) so the CSV / Parquet ingestion paths can be exercised offline.  The contents are
synthetic (class-dependent patterns), not the original datasets.

    python examples/data/make_sample_data.py --rows 2000 --out examples/data
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))

import numpy as np

from distkeras_b200.data import synthetic_higgs, synthetic_mnist

HIGGS_FEATURES = ["DER_mass_MMC", "DER_mass_transverse_met_lep", "DER_mass_vis", "DER_pt_h", "DER_deltaeta_jet_jet",
                  "DER_mass_jet_jet", "DER_prodeta_jet_jet", "DER_deltar_tau_lep", "DER_pt_tot", "DER_sum_pt",
                  "DER_pt_ratio_lep_tau", "DER_met_phi_centrality", "DER_lep_eta_centrality", "PRI_tau_pt", "PRI_tau_eta",
                  "PRI_tau_phi", "PRI_lep_pt", "PRI_lep_eta", "PRI_lep_phi", "PRI_met", "PRI_met_phi", "PRI_met_sumet",
                  "PRI_jet_num", "PRI_jet_leading_pt", "PRI_jet_leading_eta", "PRI_jet_leading_phi",
                  "PRI_jet_subleading_pt", "PRI_jet_subleading_eta", "PRI_jet_subleading_phi", "PRI_jet_all_pt"]


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=2000)
    ap.add_argument("--out", default=os.path.dirname(os.path.abspath(__file__)))
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    m = synthetic_mnist(a.rows)
    header = "label," + ",".join(f"pixel{i}" for i in range(784))
    table = np.concatenate([m["label"].numpy()[:, None].astype(np.int64), m["features"].numpy().astype(np.int64)], axis=1)
    np.savetxt(os.path.join(a.out, "mnist_sample.csv"), table, fmt="%d", delimiter=",", header=header, comments="")
    h = synthetic_higgs(a.rows)
    with open(os.path.join(a.out, "atlas_higgs_sample.csv"), "w") as f:
        f.write("EventId," + ",".join(HIGGS_FEATURES) + ",Weight,Label\n")
        x, y = h["features"].numpy(), h["label"].numpy()
        for i in range(a.rows):
            f.write(f"{100000 + i}," + ",".join(f"{v:.3f}" for v in x[i]) + f",1.0,{'s' if y[i] == 1 else 'b'}\n")
    print(f"wrote {a.rows} rows to {a.out}/mnist_sample.csv and atlas_higgs_sample.csv")


if __name__ == "__main__":
    main()
