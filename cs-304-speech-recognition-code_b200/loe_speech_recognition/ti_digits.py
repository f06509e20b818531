"""Label constants of the TIDIGITS corpus (reference: ti_digits.py:13-26).  The corpus walker
(TIDigits / DataLoader, ti_digits.py:29-203) is host I/O outside the accelerated path
 and is not part of this build."""
from typing import Dict, Literal, TypeAlias

TI_DIGITS_LABEL_TYPE: TypeAlias = Literal["1", "2", "3", "4", "5", "6", "7", "8", "9", "O", "Z"]
TI_DIGITS_LABELS: Dict[TI_DIGITS_LABEL_TYPE, int] = {
    "1": 1, "2": 2, "3": 3, "4": 4, "5": 5, "6": 6, "7": 7, "8": 8, "9": 9, "O": 0, "Z": 10,
}
