"""Mirror of `src/matching/ROC.py`: FRR over FAR, points ordered by FAR.  matplotlib is optional in this image: without
it the ordered curve is returned (and printed) instead of shown."""
from __future__ import annotations

import numpy as np


def roc_points(thresholds, far_values, frr_values):
    far, frr = np.array(far_values), np.array(frr_values)
    order = np.argsort(far)                                             # ROC.py:11-13
    return far[order], frr[order]


def plot_roc(thresholds, far_values, frr_values, title="ROC (FAR vs FRR)", show=True):
    far, frr = roc_points(thresholds, far_values, frr_values)
    try:
        import matplotlib.pyplot as plt
    except ImportError:
        print(f"{title}\n{'FAR':>8} | {'FRR':>8}")
        for a, r in zip(far, frr):
            print(f"{a:8.3f} | {r:8.3f}")
        return far, frr
    plt.figure(figsize=(7, 6))
    plt.plot(far, frr, marker="o", linewidth=2)
    plt.xlabel("FAR (False Accept Rate)")
    plt.ylabel("FRR (False Reject Rate)")
    plt.title(title)
    plt.grid(True)
    if show:
        plt.show()
    return far, frr
