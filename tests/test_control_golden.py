"""Step / layer bookkeeping of the path against the reference's own code: `LayerCounter` (which steps are full, the
odometer, its early rewind, `build_for_layer`) and `GLOBAL_CONFIG` (defaults, YAML deep merge) must behave exactly like
src/chipmunk/util/{layer_counter,config}.py.  The expected values in tests/golden/control.json were produced by running
that code unmodified (tests/golden/make_golden_control.py); where /root/reference is mounted the same comparison also
runs live, on the reference's example config files too.  CPU only."""
import copy
import glob
import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from control_cases import BUILD_CASES, CONFIG_YAMLS, COUNTER_CASES, jsonable, run_build_case, run_counter_case  # noqa: E402

REF_MOUNTED = os.path.isdir("/root/reference/src/chipmunk/util")

# Documented B200 differences of the defaults (chipmunk_b200/util/config.py): caches stay in HBM unless a
# config file switches offloading on, and an offloaded cache may live in a peer GPU's HBM.
B200_ONLY_OFFLOADING_KEYS = ("backing", "peer_device")


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(HERE, "golden", "control.json")) as f:
        return json.load(f)


def _ours_after(yaml_text, tmp_path):
    from chipmunk_b200.util import GLOBAL_CONFIG, load_from_file
    from chipmunk_b200.util.config import reset_to_defaults

    reset_to_defaults()
    p = tmp_path / "c.yml"
    p.write_text(yaml_text)
    load_from_file(str(p))
    got = jsonable(copy.deepcopy(GLOBAL_CONFIG))
    reset_to_defaults()
    return got


def _comparable(ours, yaml_sets_global_switch):
    ours = copy.deepcopy(ours)
    for k in B200_ONLY_OFFLOADING_KEYS:
        assert k in ours["offloading"], f"offloading.{k} is a documented B200 key"
        del ours["offloading"][k]
    if not yaml_sets_global_switch:
        assert ours["offloading"]["global_disable_offloading"] is True      # the B200 default: resident caches
        ours["offloading"]["global_disable_offloading"] = False             # the reference's default
    return ours


def test_default_config_matches_the_reference(cm, golden):
    from chipmunk_b200.util.config import BASE_CONFIG
    assert _comparable(jsonable(BASE_CONFIG), False) == golden["base_config"]


@pytest.mark.parametrize("name", sorted(CONFIG_YAMLS))
def test_yaml_merge_matches_the_reference(cm, golden, name, tmp_path):
    text = CONFIG_YAMLS[name]
    ours = _ours_after(text, tmp_path)
    assert _comparable(ours, "global_disable_offloading" in text) == golden["configs"][name]


@pytest.mark.parametrize("name", sorted(COUNTER_CASES))
def test_odometer_trace_matches_the_reference(cm, golden, name):
    from chipmunk_b200.util import GLOBAL_CONFIG
    from chipmunk_b200.util.config import reset_to_defaults
    from chipmunk_b200.util.layer_counter import LayerCounter

    reset_to_defaults()
    got = run_counter_case(COUNTER_CASES[name], GLOBAL_CONFIG, LayerCounter)
    reset_to_defaults()
    want = golden["traces"][name]
    assert len(got) == len(want)
    for i, (g, w) in enumerate(zip(got, want)):
        assert g == w, f"call {i}: [full_attn, full_mlp, step, layer, sub, invocation] = {g}, reference {w}"
    # the trace really contains what it is meant to pin: full and sparse steps, and the early rewind to step 0
    steps = [r[2] for r in want]
    assert any(r[0] for r in want) or COUNTER_CASES[name]["schedule"] == []
    assert any(b < a for a, b in zip(steps, steps[1:])) or COUNTER_CASES[name]["generations"] == 1


@pytest.mark.parametrize("name", sorted(BUILD_CASES))
def test_build_for_layer_matches_the_reference(cm, golden, name):
    from chipmunk_b200.util import layer_counter
    assert run_build_case(BUILD_CASES[name], layer_counter) == golden["builds"][name]


@pytest.mark.skipif(not REF_MOUNTED, reason="the reference is only mounted in the authoring container")
def test_fixture_regenerates_and_example_configs_load_alike(cm, golden, tmp_path):
    """Authoring container: control.json is what the reference's code produces today, and the reference's own example
    config files (examples/*/chipmunk-config.yml) merge into the same GLOBAL_CONFIG here as there."""
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(HERE, "golden", "make_golden_control.py"), str(tmp_path)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    with open(tmp_path / "control.json") as f:
        assert json.load(f) == golden
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden_control as gen
    files = sorted(glob.glob("/root/reference/examples/*/chipmunk-config.yml"))
    assert len(files) >= 3
    for path in files:
        text = open(path).read()
        want = jsonable(gen.reference_config_after(text, str(tmp_path)))
        ours = _ours_after(text, tmp_path)
        assert _comparable(ours, "global_disable_offloading" in text) == want, path
