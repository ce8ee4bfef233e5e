"""Drop-in check at the level the reference's users see (INTEGRATION.md option 2): the REFERENCE's own Python -- its
`SparseDiffAttn` / `SparseDiffMlp` modules, its op wrappers with their padding and slicing, its LayerCounter, config and layer
storage, staged byte-for-byte into the git-ignored oracle/_ref/chipmunk_py/ by oracle/build_ref.py -- runs unmodified on top of
this repo's `torch.ops.chipmunk.*` kernels and must reproduce the golden vectors that the same code produced over the CPU
oracle (tests/golden/modules_*.npz): stored bit masks / index sets / neuron sets exactly, every step's output and the caches
within the tolerances of tests/test_modules_golden_gpu.py.  Runs tests/ref_python_over_b200.py in a fresh interpreter, because
there `import chipmunk` must resolve to the staged reference code, not to this repo's alias package."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
STAGED = os.path.join(ROOT, "oracle", "_ref", "chipmunk_py", "chipmunk", "modules", "attn.py")


@pytest.mark.skipif(not os.path.exists(STAGED), reason="oracle/_ref/chipmunk_py not staged (python oracle/build_ref.py, authoring container)")
def test_reference_modules_run_on_the_b200_operators(cuda):
    env = dict(os.environ, TORCHDYNAMO_DISABLE="1", PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, os.path.join(HERE, "ref_python_over_b200.py")], capture_output=True, text=True,
                       env=env, cwd=ROOT, timeout=600)
    tail = (r.stdout[-3000:] + "\n" + r.stderr[-3000:])
    assert r.returncode == 0, tail
    assert "REFERENCE-PYTHON-OVER-B200 OK" in r.stdout, tail
    # all three flows ran: the two attention flows (bit-packed mask; plain index lists) and the MLP
    for marker in ("hunyuan: stored bit mask identical", "flux: stored index sets identical", "mlp: selected neuron sets identical"):
        assert marker in r.stdout, tail
