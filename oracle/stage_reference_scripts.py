"""Stage the reference's driver scripts for the GPU box (TEST INFRASTRUCTURE ONLY).

    python oracle/stage_reference_scripts.py

/root/reference does not exist on the GPU box, and reference sources are never committed to this repository.  The
acceptance test tests/test_reference_scripts.py runs ten of the reference's scripts UNMODIFIED against the drop-in
package, so it needs the script files next to it: this recipe copies them byte for byte from /root/reference/scripts into
oracle/_ref/scripts/, which is git-ignored (out of history) but not gpurun-ignored (travels with the snapshot, like a
compiled oracle/_ref library would).  __graft_entry__.build() calls it when /root/reference is present.
"""
from __future__ import annotations

import os
import shutil

SRC = "/root/reference/scripts"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "scripts")
SCRIPTS = ("project3_train.py", "project5_test_ndigits_with_sil.py", "project6_train.py",
           "project3_predict_simple.py", "project4_2digits.py", "project5_test_1digit.py", "project5_test_ndigits_no_sil.py",
           "project5_find_trans_ndigits_no_sil.py", "project5_find_trans_ndigits_with_sil.py", "project5_train_no_empty.py")


def stage() -> bool:
    if not os.path.isdir(SRC):
        return False
    os.makedirs(DST, exist_ok=True)
    for name in SCRIPTS:
        shutil.copyfile(os.path.join(SRC, name), os.path.join(DST, name))
    return True


if __name__ == "__main__":
    print("staged" if stage() else "no /root/reference here: nothing staged")
