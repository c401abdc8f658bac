"""Training history container (reference: multimodn/history.py).

The fused step fills the same fields the reference fills at multimodn.py:244-250 and :390-409:
per epoch one ``(E+1, D)`` float matrix per metric (row 0 = initial state, row e+1 = after
encoder id e) and one ``(E,)`` state-change vector.  ``get_results`` / ``save_results`` read the
last row (after the last encoder) exactly like history.py:98-161.
"""
from typing import Dict, List

import numpy as np

_METRICS = ("loss", "accuracy", "sensitivity", "specificity", "balanced_accuracy")


def display_title(key: str) -> str:
    return key.replace("_", " ").capitalize()


class MultiModNHistory:
    def __init__(self, targets: List[str]):
        self.decoder_names: List[str] = targets
        self.state_change_loss: List[np.ndarray] = []
        self.loss: Dict[str, List[np.ndarray]] = {"train": []}
        self.accuracy: Dict[str, List[np.ndarray]] = {"train": []}
        self.sensitivity: Dict[str, List[np.ndarray]] = {"train": []}
        self.specificity: Dict[str, List[np.ndarray]] = {"train": []}
        self.balanced_accuracy: Dict[str, List[np.ndarray]] = {"train": []}

    # -- writer used by MultiModN ------------------------------------------------------------
    def append(self, tag: str, matrices: Dict[str, np.ndarray]):
        for name in _METRICS:
            getattr(self, name).setdefault(tag, []).append(matrices[name])

    # -- readers (history.py:98-161) ---------------------------------------------------------
    def get_results(self):
        import pandas as pd
        names = self.decoder_names
        columns = ["State change loss"]
        cols = [[self.state_change_loss[-1][-1]] * len(names)]
        for metric in _METRICS:
            for tag, epochs in getattr(self, metric).items():
                columns.append(f"{display_title(tag)} {metric.replace('_', ' ')}")
                cols.append([epochs[-1][-1][i] for i in range(len(names))])
        df = pd.DataFrame(np.array(cols, dtype=np.float64).T, columns=columns)
        df.index = names
        return df

    def print_results(self):
        print(self.get_results())

    def save_results(self, path):
        self.get_results().to_csv(path, index_label="Target")

    def plot(self, filepath: str, targets_to_display: List[str], show_state_change: bool = False):
        """Learning curves of the last-encoder row, one subplot row per metric (history.py:34-96)."""
        import matplotlib.pyplot as plt          # optional dependency, presentation only
        tags = list(self.loss)
        fig, ax = plt.subplots(figsize=(10 * len(tags), 25), nrows=len(_METRICS), ncols=len(tags), squeeze=False)
        for name in targets_to_display:
            if name not in self.decoder_names:
                raise ValueError(f"Target name '{name}' is not part of the MultiModN history")
            i = self.decoder_names.index(name)
            for row, metric in enumerate(_METRICS):
                for col, (tag, epochs) in enumerate(getattr(self, metric).items()):
                    ax[row][col].plot([m[-1][i] for m in epochs], label=name)
                    ax[row][col].legend(loc="best")
                    ax[row][col].set_title(f"{tag.capitalize()} {display_title(metric)}")
                    ax[row][col].grid(True)
        if show_state_change:
            ax[0][0].plot([s[-1] for s in self.state_change_loss], label="State change loss")
        plt.tight_layout()
        fig.savefig(filepath)
