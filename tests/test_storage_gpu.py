"""GPU tests of the cache residency manager: resident (default), host-offloaded (the reference's
MaybeOffloadedTensor behaviour, util/storage/offloaded_tensor.py:42-178) and NVLink-peer-offloaded caches hand back
exactly what was stored, through the load_async / load_async_wait / complete_cur_layer calls the model loops make."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("backing", ["resident", "host", "peer"])
def test_attn_storage_round_trip(cm, cuda, backing):
    from chipmunk_b200.util import AttnStorage
    from chipmunk_b200.util.config import GLOBAL_CONFIG, reset_to_defaults
    reset_to_defaults()
    off = GLOBAL_CONFIG["offloading"]
    if backing != "resident":
        off["global_disable_offloading"] = False
        off["attn.out_cache"] = True
        off["attn.indices"] = True
        if backing == "peer":
            off["backing"] = "peer"
            off["peer_device"] = torch.cuda.device_count() - 1      # the neighbour GPU (the same one on a 1-GPU box)
    try:
        layers = [AttnStorage(i, init_names=["indices", "out_cache"]) for i in range(4)]
        g = torch.Generator(device=cuda).manual_seed(0)
        vals = []
        for st in layers:                                            # a "full step": every layer stores its caches
            o = torch.randn(1, 2, 500, 128, device=cuda, generator=g).to(torch.bfloat16)
            packed = torch.randint(0, 255, (4096,), device=cuda, generator=g, dtype=torch.uint8)
            st.set_out_cache(o)
            st.set_indices(packed)
            vals.append((o.clone(), packed.clone()))
            st.complete_cur_layer()
        assert layers[0].out_cache.is_offload_enabled == (backing != "resident")
        if backing == "peer":
            assert layers[0].out_cache.cpu_buf[0].device.type == "cuda"
        elif backing == "host":
            assert layers[0].out_cache.cpu_buf[0].is_pinned()
        # a "sparse step": the model loop prefetches layer i+1 while layer i computes
        layers[0].load_async()
        for i, st in enumerate(layers):
            st.load_async_wait()
            if i + 1 < len(layers):
                layers[i + 1].load_async()
            assert torch.equal(st.get_out_cache(), vals[i][0])
            assert torch.equal(st.get_indices(), vals[i][1])
            st.complete_cur_layer()
    finally:
        reset_to_defaults()
