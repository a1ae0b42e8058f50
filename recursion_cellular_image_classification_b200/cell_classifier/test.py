"""Inference + plate-group assignment — mirrors reference cell_classifier/test.py:9-58.

`test(df_test, ds_test, plate_groups, experiment_type, model, bs, num_workers, device)` keeps the reference
signature and returns the same float64 [N] array of class ids.  Softmax (test.py:27), the plate-group mask
(:42-45), the rescale (:34-39) and the greedy one-class-per-well loop (:48-56) run on the GPU
(rxb_tta_softmax_avg_mask / rxb_greedy_assign); the loop is bit-exact with numpy's (tests/golden).
`tta_views` > 1 adds the north star's D4 test-time augmentation: probabilities are averaged over the views before
masking; with the default single identity view the reference is reproduced exactly.
"""
import numpy as np
import torch

from .. import ops
from .dataloader import ImagesDS, RawView, collate_raw


def _model_logits(model, ds, batch, dev, code):
    if isinstance(model, torch.nn.DataParallel):      # main.py:94 wraps the model; one process drives one GPU here
        model = model.module
    if isinstance(ds, ImagesDS):
        b = dict(batch)
        b["codes"] = torch.full_like(batch["codes"], code)
        xs = ds.device_batch(b, dev)                                  # [B*G, H/2, W/2, 32]
        G = batch["codes"].shape[1]
        out = model(xs)                                               # [B*G, C]
        out = out.to(dev).float()
        return out.view(-1, G, out.shape[-1]).mean(1)                # site / control average (linear head)
    x, _ = batch
    return model(x.to(dev) if hasattr(model, "parameters") else x).to(dev).float()


def test(df_test, ds_test, plate_groups, experiment_type, model, bs, num_workers, device, tta_views=1):
    dev = torch.device(device)
    if isinstance(ds_test, ImagesDS):
        loader = torch.utils.data.DataLoader(RawView(ds_test), batch_size=bs, shuffle=False, num_workers=num_workers,
                                             collate_fn=collate_raw)
    else:
        loader = torch.utils.data.DataLoader(ds_test, batch_size=bs, shuffle=False, num_workers=num_workers)
    codes = [ops.aug_code(v, False, k) for v in (False, True) for k in range(4)][:max(1, tta_views)]
    views = [[] for _ in codes]
    with torch.no_grad():
        for batch in loader:
            for vi, code in enumerate(codes):
                views[vi].append(_model_logits(model, ds_test, batch, dev, code))
    logits = torch.stack([torch.cat(v, dim=0) for v in views], dim=0).contiguous()        # [V, N, C]
    assert logits.shape[1] == len(df_test)                                                # test.py:41
    plate = torch.as_tensor(np.array(df_test.plate.values), dtype=torch.int32, device=dev)
    col = torch.as_tensor(np.asarray(plate_groups[:, experiment_type]), dtype=torch.int32, device=dev)
    probs = ops.tta_softmax_avg_mask(logits, plate, col)
    return ops.greedy_assign(probs).cpu().numpy().astype(np.float64)
