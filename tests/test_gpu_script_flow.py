"""GPU: the reference's driver workflow (silence stripping, training, saving, loading, penalty poke,
ProcessPoolExecutor over predict) through the drop-in package, as a separate process like a script."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_project5_workflow_runs_like_a_script(built_lib, tmp_path):
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "examples", "digits_pipeline.py"), str(tmp_path)],
                                  text=True, timeout=600)
    assert "through a process pool" in out
    acc = float(out.split("accuracy")[1].split(";")[0])
    assert acc >= 0.5
    assert (tmp_path / "big_model_speech_only" / "S" / "multivariate_normals.pickle").exists()
