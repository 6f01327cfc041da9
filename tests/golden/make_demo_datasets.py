"""Training sets and starting inducing points of the reference's FOUR demos outside BASELINE.json's configs, produced by
the reference's OWN loaders (utils/dataset_utils.py, run unmodified through oracle/run_reference.py) and by
scipy.cluster.vq.kmeans exactly as the demos call it (seeds 0 / 1) — so that examples/replay_demos.py can replay them as
full training runs on the GPU box, which has neither /root/reference nor the CSV.

Run in the authoring container only:   python tests/golden/make_demo_datasets.py
Writes tests/golden/datasets/demo_datasets.npz (a few tens of KB)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import run_reference as ref  # noqa: E402


def main():
    from scipy.cluster.vq import kmeans
    out = {}
    for demo, dataset in (("tf2_2d", "toy_2d"),                      # demos/demo_tf2_2d.py:22
                          ("tf2_modified", "toy_multimodal"),        # demos/demo_tf2_modified.py:22
                          ("tf2_modified_multiclass", "toy_categorical"),   # demos/demo_tf2_modified_multiclass.py:22
                          ("john_doe_multi_class", "john_doe_boundary")):   # demos/demo_john_doe_multi_class.py:23
        N, Xtr, Ytr, _ = ref.load_reference_dataset(dataset, 0)
        Xtr, Ytr = np.asarray(Xtr, dtype=np.float64), np.asarray(Ytr, dtype=np.float64).reshape(-1, 1)
        out[f"{demo}.X"], out[f"{demo}.Y"] = Xtr, Ytr
        out[f"{demo}.Z"], out[f"{demo}.Z_assign"] = kmeans(Xtr, 25, seed=0)[0], kmeans(Xtr, 25, seed=1)[0]   # e.g. demo_tf2_2d.py:39
        print(demo, Xtr.shape, Ytr.shape, out[f"{demo}.Z"].shape, out[f"{demo}.Z_assign"].shape, "targets", np.unique(Ytr)[:6])
    np.savez_compressed(os.path.join(HERE, "datasets", "demo_datasets.npz"), **out)


if __name__ == "__main__":
    main()
