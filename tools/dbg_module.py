import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chipmunk_b200 as cm
from oracle import chipmunk_oracle as oracle
from chipmunk_b200.util.config import reset_to_defaults
from chipmunk_b200.util import layer_counter as lc
BF = torch.bfloat16
cuda = torch.device("cuda", 0)
for tsel in (False, True):
    reset_to_defaults()
    cfg = cm.util.GLOBAL_CONFIG
    cfg["steps"] = 50
    cfg["attn"].update(first_n_dense_layers=0, top_keys=0.3, recompute_mask=False, should_compress_indices=False,
                       pad_qkv_before_kernel=False, counts_multiple_of=112, random_keys=0.0, torch_selection=tsel)
    lc.singleton.__init__(0, 0)
    layer_num, counter = cm.LayerCounter.build_for_layer(is_attn_sparse=True)
    attn = cm.SparseDiffAttn(layer_num, counter)
    g = torch.Generator(device=cuda).manual_seed(0)
    B, H, N = 1, 2, 1000
    q, k, v = (torch.randn(B, H, N, 128, device=cuda, generator=g).to(BF) for _ in range(3))
    outs = [attn(q, k, v) for _ in range(3)]
    q2 = (q.float() + 0.3 * torch.randn(q.shape, device=cuda, generator=g)).to(BF)
    cache_before = attn.storage.get_out_cache().clone()
    inds, counts = attn.storage.get_indices(), attn.storage.get_counts()
    print("torch_sel", tsel, "inds", tuple(inds.shape), inds.dtype, "counts", counts.flatten().tolist()[:8], "step", counter.cur_inference_step)
    o3 = attn(q2, k, v)
    ref = oracle.csp_attn(q2.cpu(), k.cpu(), v.cpu(), cache_before.cpu(), inds.cpu(), counts.cpu(), 1)
    rel = ((o3.cpu().float() - ref.float()).norm() / ref.float().norm()).item()
    print("rel o3 vs oracle", rel)
    d = (o3.cpu().float() - ref.float()).abs().amax(dim=-1)   # per row
    print("worst rows", d.flatten().topk(5))
    for h in range(H):
        for gi in range(inds.shape[2]):
            c = int(counts[0, h, gi]); s = inds[0, h, gi, :c]
            print(h, gi, c, "min", int(s.min()), "max", int(s.max()), "unique", int(s.unique().numel()))
