"""Inference + plate-group assignment — mirrors reference cell_classifier/test.py:9-58.

`test(df_test, ds_test, plate_groups, experiment_type, model, bs, num_workers, device)` keeps the reference
signature and returns the same float64 [N] array of class ids.  Softmax (test.py:27), the plate-group mask
(:42-45), the rescale (:34-39) and the greedy one-class-per-well loop (:48-56) run on the GPU
(rxb_tta_softmax_avg_mask / rxb_greedy_assign); the loop is bit-exact with numpy's (tests/golden).
`tta_views` > 1 adds the north star's D4 test-time augmentation: probabilities are averaged over the views before
masking; with the default single identity view the reference is reproduced exactly.

Which images feed a prediction: the sample's own two sites, averaged (the first third of the reference's item,
models.py:46-49; see cell_classifier/models.py) — the same rule train() and evaluate() use.  Control wells are not
decoded on this path.

Under torchrun (WORLD_SIZE > 1) the wells are sharded over the ranks (contiguous ranges), every rank runs its shard,
the [N, C] logits are all-gathered and every rank computes the same assignment (the greedy loop is
sequential over the N picks; the 18 test experiments are independent replicas).
"""
import numpy as np
import torch

from .. import ops, parallel
from .dataloader import ImagesDS, RawView, collate_raw
from .models import sample_group


def _model_logits(model, ds, batch, dev, code):
    if isinstance(model, torch.nn.DataParallel):      # main.py:94 wraps the model; one process drives one GPU here
        model = model.module
    if isinstance(ds, ImagesDS):
        b = dict(batch)
        b["codes"] = torch.full_like(batch["codes"], code)
        xs = ds.device_batch(b, dev)                                  # [B*G, H/2, W/2, 32]
        G = batch["codes"].shape[1]
        if getattr(model, "wants_controls", False):                   # the reference's own model: thirds concatenated
            return model(xs, G=G).to(dev).float()                     # inside (models.py:44-55)
        assert sample_group(G) == G, "test() loads items without control wells"
        out = model(xs)                                               # [B*G, C], G = the sample's sites
        out = out.to(dev).float()
        return out.view(-1, G, out.shape[-1]).mean(1)                # site average (linear head: = feature average)
    x, _ = batch
    return model(x.to(dev) if hasattr(model, "parameters") else x).to(dev).float()


def predict_probs(df_test, ds_test, plate_groups, experiment_type, model, bs, num_workers, device, tta_views=1):
    """test.py:18-46: the [N, C] float32 matrix the greedy loop starts from — softmax per view, mean over the D4
    views, plate-group mask, rescale — as a device tensor (identical on every rank)."""
    dev = torch.device(device)
    rank, world = parallel.rank_world()
    net = model.module if isinstance(model, torch.nn.DataParallel) else model
    view = RawView(ds_test, controls=bool(getattr(net, "wants_controls", False))) if isinstance(ds_test, ImagesDS) else ds_test
    counts = None
    if world > 1:                                                     # shard the wells: contiguous range per rank
        counts = [e - b for b, e in (parallel.shard_range(len(view), r, world) for r in range(world))]
        b, e = parallel.shard_range(len(view), rank, world)
        view = torch.utils.data.Subset(view, range(b, e))
    kw = {"collate_fn": collate_raw} if isinstance(ds_test, ImagesDS) else {}
    loader = torch.utils.data.DataLoader(view, batch_size=bs, shuffle=False, num_workers=num_workers, **kw)
    codes = [ops.aug_code(v, False, k) for v in (False, True) for k in range(4)][:max(1, tta_views)]
    views = [[] for _ in codes]
    with torch.no_grad():
        for batch in loader:
            for vi, code in enumerate(codes):
                views[vi].append(_model_logits(model, ds_test, batch, dev, code))
    if counts is not None:                                            # a rank with an empty shard still joins the gather
        C = parallel.max_int(views[0][0].shape[-1] if views[0] else 0, dev)
        if not views[0]:
            views = [[torch.zeros(0, C, dtype=torch.float32, device=dev)] for _ in codes]
    logits = torch.stack([torch.cat(v, dim=0) for v in views], dim=0).contiguous()        # [V, n_local, C]
    if counts is not None:
        V, _, C = logits.shape
        rows = parallel.allgather_rows(logits.permute(1, 0, 2).reshape(-1, V * C).contiguous(), counts)
        logits = rows.view(-1, V, C).permute(1, 0, 2).contiguous()                        # [V, N, C] on every rank
    assert logits.shape[1] == len(df_test)                                                # test.py:41
    plate = torch.as_tensor(np.array(df_test.plate.values), dtype=torch.int32, device=dev)
    col = torch.as_tensor(np.asarray(plate_groups[:, experiment_type]), dtype=torch.int32, device=dev)
    return ops.tta_softmax_avg_mask(logits, plate, col)


def test(df_test, ds_test, plate_groups, experiment_type, model, bs, num_workers, device, tta_views=1):
    probs = predict_probs(df_test, ds_test, plate_groups, experiment_type, model, bs, num_workers, device, tta_views)
    return ops.greedy_assign(probs).cpu().numpy().astype(np.float64)
