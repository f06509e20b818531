"""Result plots of the reference's test drivers (visualizer.py:6-67): a confusion matrix and a line plot,
saved under ``./plots`` (the directory must exist, as in the reference).  matplotlib is imported when a plot is
made, not at package import, so the decoder does not depend on it."""
from __future__ import annotations

import numpy as np


def _pyplot():
    import matplotlib
    matplotlib.use("Agg", force=False)
    import matplotlib.pyplot as plt
    return plt


def confusion_counts(predictions, ground_truth, class_names) -> np.ndarray:
    """counts[true, predicted] over the label lists (ValueError for a label outside ``class_names``)."""
    names = list(class_names)
    counts = np.zeros((len(names), len(names)), dtype=int)
    for truth, predicted in zip(ground_truth, predictions):
        counts[names.index(truth), names.index(predicted)] += 1
    return counts


def plot_confusion_matrix_from_lists(predictions, ground_truth, class_names, title="Confusion Matrix", figsize=(8, 6)):
    plt = _pyplot()
    counts = confusion_counts(predictions, ground_truth, class_names)
    plt.figure(figsize=figsize)
    plt.imshow(counts, interpolation="nearest")
    plt.title(title)
    plt.colorbar()
    ticks = np.arange(len(class_names))
    plt.xticks(ticks, class_names, rotation=45)
    plt.yticks(ticks, class_names)
    half = counts.max() / 2.0
    for i, j in np.ndindex(counts.shape):
        plt.text(j, i, format(counts[i, j], "d"), ha="center", va="center", color="white" if counts[i, j] > half else "black")
    plt.tight_layout()
    plt.ylabel("True label")
    plt.xlabel("Predicted label")
    plt.savefig(f"./plots/confusion_matrix_{title}.png")


def plot_line(x_values, y_values, title="Line Plot", x_label="X-axis", y_label="Y-axis"):
    if len(x_values) != len(y_values):
        raise ValueError("The lengths of x_values and y_values must be the same.")
    plt = _pyplot()
    plt.plot(x_values, y_values)
    plt.title(title)
    plt.xlabel(x_label)
    plt.ylabel(y_label)
    plt.grid(True)
    plt.savefig("./plots/" + title.replace(" ", "_") + ".png")
