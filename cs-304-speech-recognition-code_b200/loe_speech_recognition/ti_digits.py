"""TIDIGITS corpus access (reference: ti_digits.py:13-203): label constants, the lazy per-label
``DataLoader`` and the ``TIDigits`` walker over ``<root>/{Adults,Children}/TIDIGITS/{TRAIN,TEST}/**.wav``.

Same classes, methods and quirks as the reference (label = file name up to the first dot minus its last
character, ti_digits.py:125-129; ``+`` extends the left operand's lists in place, :44-51; a missing folder
yields an empty loader because ``os.walk`` is silent, :90-123).  One addition for SURVEY §8 f1:
``DataLoader.sample_dtype``.  The reference converts every WAV to float32 on the host (:137-139);
setting ``DataLoader.sample_dtype = np.int16`` keeps 16-bit PCM files as int16, which ``MFCC.batch`` and the
``decode_pcm_*`` entry points accept as such -- half the PCIe bytes, bit-identical features.
"""
from __future__ import annotations

import logging
import os
from dataclasses import dataclass, field
from typing import Any, Dict, Generator, List, Literal, Tuple, TypeAlias, Union

import numpy as np
from numpy.typing import NDArray

logger = logging.getLogger(__name__)

TI_DIGITS_LABEL_TYPE: TypeAlias = Literal["1", "2", "3", "4", "5", "6", "7", "8", "9", "O", "Z"]
TI_DIGITS_LABELS: Dict[TI_DIGITS_LABEL_TYPE, int] = {
    "1": 1, "2": 2, "3": 3, "4": 4, "5": 5, "6": 6, "7": 7, "8": 8, "9": 9, "O": 0, "Z": 10,
}


@dataclass
class DataLoader:
    data: Dict[str, List[Union[NDArray, str]]]

    # dtype handed to callers: float32 like the reference, or int16 to keep 16-bit PCM narrow (see module docstring)
    sample_dtype = np.float32

    def __post_init__(self) -> None:
        logger.info("Create TIDigits Data Loader with %d labels", len(self))

    def __len__(self) -> int:
        return len(self.data)

    def __iter__(self) -> Generator[Tuple[NDArray, str], Any, Any]:
        for label, clips in self.data.items():
            for clip in clips:
                yield self.lazy_loading(clip), label

    def __add__(self, other: "DataLoader") -> "DataLoader":
        merged = self.data                       # the reference merges into the left operand's dict (no copy)
        for label, clips in other.data.items():
            if label in merged:
                merged[label].extend(clips)
            else:
                merged[label] = clips
        return type(self)(merged)

    def __getitem__(self, key: str) -> List[NDArray]:
        """All clips of one label, loaded."""
        return [self.lazy_loading(clip) for clip in self.data[key]]

    def get_combined(self, labels: str, key: int = 0) -> NDArray:
        """Clip number ``key`` of every label character in ``labels``, concatenated (a synthetic digit string)."""
        return np.concatenate([self[label][key] for label in labels])

    def get_all_n_digits(self, n: int) -> Dict[str, List[NDArray]]:
        return {label: [self.lazy_loading(clip) for clip in clips]
                for label, clips in self.data.items() if len(label) == n}

    @classmethod
    def from_folder_path(cls, folder_path: str, isLazyLoading: bool = True) -> "DataLoader":
        data: Dict[str, List[Union[NDArray, str]]] = {}
        for dirpath, _dirnames, filenames in os.walk(folder_path):
            for filename in filenames:
                if not (filename.endswith(".wav") or filename.endswith(".WAV")):
                    continue
                path = os.path.join(dirpath, filename)
                label = cls.filename_parser(filename)
                data.setdefault(label, []).append(path if isLazyLoading else cls.lazy_loading(path))
        return cls(data)

    @staticmethod
    def filename_parser(file_name: str) -> str:
        """'12a.wav' -> '12': everything before the first dot, minus the production letter."""
        return file_name.split(".")[0][:-1]

    @classmethod
    def lazy_loading(cls, clip: Union[str, NDArray]) -> NDArray:
        if isinstance(clip, np.ndarray):
            return clip
        if isinstance(clip, str):
            from scipy.io import wavfile
            samples = wavfile.read(clip)[1]
            if cls.sample_dtype == np.int16 and samples.dtype == np.int16:
                return samples
            return samples.astype(np.float32)
        raise NotImplementedError(f"Cannot deal with {type(clip)}")


@dataclass
class TIDigits:
    folder_path: str

    include_adult: bool = field(default=True)
    include_children: bool = field(default=True)
    include_percentage: float = field(default=1.0)      # accepted and ignored, like the reference
    isLazyLoading: bool = field(default=True)

    _train_dataset: DataLoader = field(init=False)
    _test_dataset: DataLoader = field(init=False)

    def __post_init__(self) -> None:
        self._train_dataset = DataLoader({})
        self._test_dataset = DataLoader({})
        groups = (("Adults", self.include_adult), ("Children", self.include_children))
        for group, wanted in groups:
            if not wanted:
                continue
            base = os.path.join(self.folder_path, group, "TIDIGITS")
            self._train_dataset += DataLoader.from_folder_path(os.path.join(base, "TRAIN"), self.isLazyLoading)
            self._test_dataset += DataLoader.from_folder_path(os.path.join(base, "TEST"), self.isLazyLoading)
        if not self.include_adult and not self.include_children:
            logger.error("Both Adults and Children are not included")
            raise Exception
        logger.info("Successfully create TIDigits dataset")

    @property
    def train_dataset(self) -> DataLoader:
        return self._train_dataset

    @property
    def test_dataset(self) -> DataLoader:
        return self._test_dataset
