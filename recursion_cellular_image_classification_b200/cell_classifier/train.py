"""Training loop of the hot path — mirrors reference cell_classifier/train.py:18-141 without ignite.

`train(experiment_id, ds_train, ds_val, model, optimizer, hyperparams, num_workers, device, debug)` keeps the
reference signature and side effects: per batch zero_grad/forward/CrossEntropy/backward/step (train.py:37,44),
an evaluation before the first epoch and after every epoch with Loss and Accuracy (:39-42,82-102), the best
validation accuracy's state_dict saved to models/best_model_<id>.pth with DataParallel's `module.` key prefix
(:88-96, main.py:147), cosine annealing per epoch with eta_min = lr/100 (:104-112) and optional early stopping
(:74-80), and with hyperparams['pretrained'] the two-epoch freeze of everything but the head (:46-67).
TensorBoard scalars go to board/<id> with the reference's tags (:114-135); gradient histograms and the progress bar
(:69-70,136-138) are not reproduced.

The step itself runs natively: workers decode JPEGs, the u8 batch goes to the GPU, the fused loader normalises
and augments it into the stem conv's layout, the DenseNet-121 executor does forward / loss / backward, gradients
are all-reduced over NCCL in phases when several ranks train, and the fused nesterov-SGD kernel updates the flat
parameter buffer.  `optimizer` is read for its hyper-parameters (lr, momentum, nesterov, weight_decay).
"""
import math
import os
import time

import torch

from .. import ops, parallel
from .dataloader import RawView, collate_raw


def _unwrap(model):
    return model.module if hasattr(model, "module") else model


def cosine_lr(lr0, epoch, nb_epochs):
    """torch CosineAnnealingLR(T_max=nb_epochs, eta_min=lr0/100) evaluated after `epoch` scheduler steps."""
    eta_min = lr0 / 100
    return eta_min + (lr0 - eta_min) * (1 + math.cos(math.pi * epoch / nb_epochs)) / 2


def evaluate(model, ds, bs, num_workers, device):
    """Loss and Accuracy over a dataset (train.py:39-42) in eval mode; returns (accuracy, mean loss)."""
    net = _unwrap(model)
    was_training = net.training
    net.eval()
    triplet = bool(getattr(net, "wants_controls", False))       # the reference's own model consumes the control thirds
    loader = torch.utils.data.DataLoader(RawView(ds, controls=triplet), batch_size=bs, shuffle=False,
                                         num_workers=num_workers, collate_fn=collate_raw)
    correct, total, loss_sum = 0, 0, 0.0
    dev = torch.device(device)
    for batch in loader:
        xs = ds.device_batch(batch, dev)                              # one image per sample (its own site, no controls)
        y = torch.tensor(batch["label"], dtype=torch.int64, device=dev)
        logits = net(xs, G=batch["codes"].shape[1]) if triplet else net(xs)
        loss_rows, _ = ops.softmax_ce(logits, y)
        loss_sum += loss_rows.sum().item()
        correct += (logits.argmax(1) == y).sum().item()
        total += y.numel()
    net.train(was_training)
    return (correct / max(total, 1)), (loss_sum / max(total, 1))


def train(experiment_id, ds_train, ds_val, model, optimizer, hyperparams, num_workers, device, debug=False):
    net = _unwrap(model)
    dev = torch.device(device)
    rank, _, world = parallel.init_from_env()
    bs = hyperparams['bs'] // world if world > 1 else hyperparams['bs']      # hyperparams['bs'] is the global batch
    group = optimizer.param_groups[0]
    lr0 = hyperparams.get('lr', group['lr'])
    momentum, nesterov, wd = group.get('momentum', 0.0), group.get('nesterov', False), group.get('weight_decay', 0.0)
    nb_epochs = hyperparams['nb_epochs']
    # TwoSitesResNet50 (the reference's own model): items keep their control wells, one native step per batch
    # (models.TwoSitesResNet50.train_step), default crop 364 like the reference (dataloader.py:47,50)
    triplet = bool(getattr(net, "wants_controls", False))
    crop = hyperparams.get('crop', 364 if triplet else 512)
    ds_train.crop = ds_val.crop = crop

    sampler = None
    if world > 1:
        sampler = torch.utils.data.distributed.DistributedSampler(ds_train, num_replicas=world, rank=rank, shuffle=True)
    # items carry the sample's own image only (controls=False): the control wells cannot reach DenseNet's single
    # linear classifier (see cell_classifier/models.py), so they are neither decoded nor copied — train, validation
    # and test() all follow this one rule
    loader = torch.utils.data.DataLoader(RawView(ds_train, controls=triplet), batch_size=bs, shuffle=sampler is None,
                                         sampler=sampler, num_workers=num_workers, collate_fn=collate_raw,
                                         drop_last=world > 1)
    net.train()
    # TensorBoard scalars under board/<id> like train.py:114-138 (tags 'training/loss' and 'lr/group_0' per iteration,
    # 'validation/accuracy' and 'validation/loss' per epoch); gradient histograms (:136-138) are not written.
    writer = None
    if rank == 0 and hyperparams.get('tensorboard', True):
        try:
            from torch.utils.tensorboard import SummaryWriter
            writer = SummaryWriter('board/' + experiment_id)
        except Exception:                                                     # logging is optional, compute is not
            writer = None
    iteration = 0
    best_acc, best_epoch, history = -1.0, 0, []
    loss_dev = torch.zeros(1, device=dev)
    n_phases = None
    # CUDA-graph replay of the executor's phases (models.DenseNet121.train_step): on by default on the device;
    # hyperparams['cuda_graph'] = False or RXB_NO_GRAPH=1 keeps plain stream launches
    use_graph = bool(hyperparams.get('cuda_graph', True)) and dev.type == "cuda" and hasattr(net, "static_buffers") \
        and os.environ.get("RXB_NO_GRAPH", "0") != "1" and not triplet
    graph_kw = {"graph": True} if use_graph else {}

    def validate(epoch):
        nonlocal best_acc, best_epoch
        acc, loss = evaluate(model, ds_val, bs, num_workers, device)
        if world > 1:
            # every rank evaluates, but validation items draw a random site per image (dataloader.py:159-173), so the
            # ranks' numbers differ slightly: rank 0's decide (best checkpoint, early stopping) for everybody
            m = torch.tensor([acc, loss], dtype=torch.float64, device=dev)
            torch.distributed.broadcast(m, src=0)
            acc, loss = float(m[0]), float(m[1])
        history.append({"epoch": epoch, "val_acc": acc, "val_loss": loss})
        if writer is not None:
            writer.add_scalar('validation/accuracy', acc, epoch)
            writer.add_scalar('validation/loss', loss, epoch)
        if rank == 0:
            print("Validation Results - Epoch: {}  Average accuracy: {:.4f} Average loss: {:.4f}".format(epoch, acc, loss))
        if acc > best_acc:                                                    # train.py:88-96
            best_acc, best_epoch = acc, epoch
            if rank == 0:
                os.makedirs('models', exist_ok=True)
                sd = {"module." + k: v.cpu() for k, v in net.state_dict().items()}
                torch.save(sd, 'models/best_model_' + experiment_id + '.pth')
        return acc

    validate(0)                                                               # Events.STARTED evaluation (train.py:82)
    for epoch in range(1, nb_epochs + 1):
        lr = cosine_lr(lr0, epoch - 1, nb_epochs) if hyperparams.get('scheduler', True) else lr0
        # train.py:46-67: a pretrained trunk stays frozen for the first two epochs (only the head learns)
        frozen = {"head_only": True} if hyperparams.get('pretrained', False) and epoch < 3 else {}
        if rank == 0 and hyperparams.get('pretrained', False) and epoch in (1, 3):
            print('classifier is unfrozen' if epoch == 1 else 'Turn on all the layers')
        if sampler is not None:
            sampler.set_epoch(epoch)
        loss_hist = torch.zeros(max(len(loader), 1), device=dev)             # per-iteration losses, read back once per epoch
        n_it, n_img = 0, 0
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)
        t_epoch = time.perf_counter()
        for batch in loader:
            y = torch.tensor(batch["label"], dtype=torch.int64, device=dev)
            if triplet:
                xs = ds_train.device_batch(batch, dev)                        # [B*3, h/2, w/2, 32]: image, neg, pos per sample
                G = batch["codes"].shape[1]
                B, H, W = xs.shape[0] // G, xs.shape[1] * 2, xs.shape[2] * 2
                net.train_step(xs, y, G=G, global_batch=B * world, loss_out=loss_dev)
                if world > 1:
                    torch.distributed.all_reduce(net.flat.grad)
                net.sgd_step(B, G, H, W, lr=lr, momentum=momentum, weight_decay=wd, nesterov=nesterov, **frozen)
                loss_hist[n_it:n_it + 1].copy_(loss_dev.to(loss_hist.dtype))
                n_it += 1
                n_img += B * G
                continue
            if use_graph:
                # the loader writes straight into the executor's fixed-address step buffer, the labels are copied
                # beside it, and every backward phase is replayed from a CUDA graph (captured on its second use)
                B, H, W = len(batch["label"]), batch["out"], batch["out"]
                xs_static, y_static, _ = net.static_buffers(B, H, W)
                xs = ds_train.device_batch(batch, dev, out=xs_static)
                y_static.copy_(y)
                y = y_static
            else:
                xs = ds_train.device_batch(batch, dev)
            B, H, W = xs.shape[0], xs.shape[1] * 2, xs.shape[2] * 2
            if n_phases is None:
                from .._lib import load
                n_phases = load().rxb_dn121_num_phases()
            ranges = [net.phase_grad_range(B, H, W, p) for p in range(n_phases)]
            ar = parallel.PhasedGradAllReduce(net.flat.grad, ranges)
            for p in range(n_phases):
                net.train_step(xs, y, global_batch=B * world, phase=p, loss_out=loss_dev, **graph_kw)
                ar.after_phase(p)
            ar.wait()
            net.sgd_step(B, H, W, lr=lr, momentum=momentum, weight_decay=wd, nesterov=nesterov, **frozen)
            loss_hist[n_it:n_it + 1].copy_(loss_dev.to(loss_hist.dtype))
            n_it += 1
            n_img += B
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)
        t_epoch = time.perf_counter() - t_epoch
        if world > 1:
            torch.distributed.all_reduce(loss_hist)                           # the ranks' shares add up to the batch mean
        if writer is not None:
            for i, v in enumerate(loss_hist[:n_it].tolist()):
                writer.add_scalar('training/loss', v, iteration + i + 1)
                writer.add_scalar('lr/group_0', lr, iteration + i + 1)
        iteration += n_it
        validate(epoch)
        # throughput of the epoch's training loop through the DataLoader (this rank's images; wall clock, synchronised)
        history[-1].update({"train_seconds": t_epoch, "train_images": n_img, "train_loss_last": float(loss_hist[max(n_it - 1, 0)])})
        if hyperparams.get('early_stopping', False) and epoch - best_epoch >= hyperparams.get('patience', 10):
            break
    if writer is not None:
        writer.close()
    return history
