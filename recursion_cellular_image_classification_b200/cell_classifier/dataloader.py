"""Dataset of the hot path — mirrors reference cell_classifier/dataloader.py.

`ImagesDS` keeps the reference constructor (dataloader.py:17-24), the RAM cache of raw JPEG bytes keyed
[experiment][plate][well] -> [site1[6], site2[6]] (:75-109), the negative-control (well B02) and random
positive-control lookup (:161-173, 188-203) and `__getitem__`'s return types (:177-180, 207-209).
What changes is WHERE the arithmetic runs: JPEG decode stays on the host (cv2, :141-146) but flip / rotate /
crop / normalize (:128-139) run in the fused GPU loader (rxb_load_norm_aug), with the augmentation drawn
explicitly (albumentations' RNG stream is not reproducible):
  vflip, hflip ~ Bernoulli(.5) (:43-44); rotation k*90 degrees, k ~ U{0..3} — the D4 subset of
  ShiftScaleRotate(rotate_limit=180) (:45-46); crop offsets ~ U{0..H-crop} (RandomCrop, :47) or centred (:50).
  `decode='gpu'` also moves the JPEG decode (:141-146) to the device (rxb_jpeg_decode_gray, bit-identical to
  cv2.imdecode): workers then only pick byte strings from the RAM cache.
  `augment='rotate'` restores the reference's full train transform: angle ~ U(-180, 180) about (w/2, h/2), bilinear,
  BORDER_REFLECT_101, through rxb_load_norm_affine (u8 result bit-identical to cv2.warpAffine).

Two ways to consume it:
  * `ds[i]` — reference-compatible: a float32 tensor [3,6,h,w] (train/val) or [6,6,H,W] (test) and the label /
    id_code.  The tensor is produced by the GPU kernel in its bit-exact fp32 mode, so use num_workers=0.
  * `ds.raw_item(i)` + `collate_raw` + `device_batch` — the fast path used by train()/test(): workers only decode;
    the batch travels as u8 (1.5 MB/image instead of 6.3 MB fp32) and the loader kernel writes the stem conv's
    bf16 layout directly.
"""
import random
from copy import deepcopy

import numpy as np
import torch

from .. import ops


class ImagesDS(torch.utils.data.Dataset):
    def __init__(self, df, df_controls, stats_experiments, img_dir, mode, verbose=True,
                 channels=[1, 2, 3, 4, 5, 6], crop=364, device="cuda", augment="d4", decode="host"):
        if augment not in ("d4", "rotate"):
            raise ValueError("augment must be 'd4' or 'rotate'")
        if decode not in ("host", "gpu"):
            raise ValueError("decode must be 'host' or 'gpu'")
        self.augment = augment
        self.decode = decode
        self.records = deepcopy(df).to_records(index=False)
        df_conts = deepcopy(df_controls)
        mask = (df_conts['well_type'] == 'negative_control') & (df_conts['well'] == 'B02')
        self.records_neg_conts = df_conts[mask].to_records(index=False)
        mask = (df_conts['well_type'] == 'positive_control')
        self.records_pos_conts = df_conts[mask].to_records(index=False)
        self.stats_exps = stats_experiments
        self.mode = mode
        self.channels = channels
        self.img_dir = img_dir
        self.len = df.shape[0]
        self.crop = crop
        self.device = device
        self.experiments = sorted(stats_experiments.keys())
        self.exp_index = {e: i for i, e in enumerate(self.experiments)}
        mean = np.stack([np.asarray(stats_experiments[e]['mean'], dtype=np.float64) for e in self.experiments])
        std = np.stack([np.asarray(stats_experiments[e]['std'], dtype=np.float64) for e in self.experiments])
        self.norm_m, self.norm_d = ops.normalize_constants(mean, std)     # float32 [n_exp, 6]
        self._norm_dev = None
        self.imgs = self._load_imgs(self.records)
        self.imgs_neg_conts = self._load_imgs(self.records_neg_conts)
        self.imgs_pos_conts = self._load_imgs(self.records_pos_conts)

    # ------------------------------------------------------------ index + byte cache (dataloader.py:64-109)
    def _get_img_path(self, records, index, channel, site):
        exp, plate, well = records[index].experiment, records[index].plate, records[index].well
        mode = 'train' if self.mode in ('train', 'val') else 'test'
        return '/'.join([self.img_dir, mode, exp, f'Plate{plate}', f'{well}_s{site}_w{channel}.jpeg'])

    def _load_imgs(self, records):
        imgs_dict = dict()
        for index in range(len(records)):
            sites = []
            for site in (1, 2):
                bufs = []
                for ch in self.channels:
                    with open(self._get_img_path(records, index, ch, site), 'rb') as f:
                        bufs.append(f.read())
                sites.append(bufs)
            exp, plate, well = records[index].experiment, records[index].plate, records[index].well
            imgs_dict.setdefault(exp, dict()).setdefault(plate, dict())[well] = sites
        return imgs_dict

    def _load_from_buffer(self, img_buffer):
        import cv2
        return np.stack([cv2.imdecode(np.frombuffer(b, dtype=np.uint8), -1) for b in img_buffer])   # u8 [6,H,W]

    # ------------------------------------------------------------ augmentation draw (explicit parameters)
    def _draw(self, S):
        """(aug code, crop offset, output size, forward warp matrix or None) for one image."""
        if self.mode == 'train':
            vflip, hflip = random.random() < 0.5, random.random() < 0.5
            if self.augment == 'rotate':
                code, M = ops.aug_code(vflip, hflip), ops.rotation_matrix(S, S, random.uniform(-180, 180))
            else:
                code, M = ops.aug_code(vflip, hflip, random.randint(0, 3)), None
            c = self.crop
            return code, (int((S - c) * random.random()), int((S - c) * random.random())), c, M
        if self.mode == 'val':
            c = self.crop
            return 0, ((S - c) // 2, (S - c) // 2), c, None
        return 0, (0, 0), S, None

    def raw_item(self, index, controls=True):
        """Decoded u8 planes and augmentation parameters; no arithmetic on the host.
        controls=True: the reference's item (dataloader.py:148-209) — the sample's image(s) followed by the plate's
        negative- and positive-control image(s): G = 3 (train/val) or 6 (test).  controls=False: the sample's own
        image(s) only, G = 1 or 2 — what a single-image trunk with a linear head consumes (DenseNet121: the control
        thirds never reach its classifier), so they are neither decoded nor copied to the device."""
        rec = self.records[index]
        exp, plate, well = rec.experiment, rec.plate, rec.well
        if self.mode in ('train', 'val'):
            picks = [self.imgs[exp][plate][well][random.randint(0, 1)]]
            if controls:
                pos_wells = list(self.imgs_pos_conts[exp][plate].keys())
                picks += [self.imgs_neg_conts[exp][plate]['B02'][random.randint(0, 1)],
                          self.imgs_pos_conts[exp][plate][random.sample(pos_wells, 1)[0]][random.randint(0, 1)]]
            label = int(rec.sirna)
        else:
            picks = list(self.imgs[exp][plate][well])
            if controls:
                pos_wells = list(self.imgs_pos_conts[exp][plate].keys())
                pos = self.imgs_pos_conts[exp][plate][random.sample(pos_wells, 1)[0]]
                picks += list(self.imgs_neg_conts[exp][plate]['B02']) + list(pos)
            label = rec.id_code
        if self.decode == 'gpu':
            planes = None
            S = ops.jpeg_frame_size(picks[0][0])[1]
        else:
            planes = np.stack([self._load_from_buffer(p) for p in picks])        # [G,6,H,W] u8
            S = planes.shape[-1]
        draws = [self._draw(S) for _ in picks]                                   # independent per image (:159-173)
        codes = np.array([d[0] for d in draws], dtype=np.uint8)
        crops = np.array([d[1] for d in draws], dtype=np.int32)
        item = {"planes": torch.from_numpy(planes) if planes is not None else None, "codes": torch.from_numpy(codes), "crops": torch.from_numpy(crops),
                "exp": self.exp_index[exp], "out": draws[0][2], "label": label}
        if planes is None:
            del item["planes"]
            item["jpeg"] = [b for p in picks for b in p]                         # G*6 byte strings, channel-minor
            item["size"] = S
        if draws[0][3] is not None:
            item["mats"] = torch.from_numpy(np.stack([d[3] for d in draws]))     # [G,2,3] float64
        return item

    def _norm(self, dev):
        if self._norm_dev is None or self._norm_dev[0].device != dev:
            self._norm_dev = (torch.from_numpy(self.norm_m).to(dev), torch.from_numpy(self.norm_d).to(dev))
        return self._norm_dev

    def device_batch(self, batch, dev, out_format=ops.OUT_BF16_S2D32, first_only=False, out=None):
        """collate_raw output -> normalised/augmented device tensor via the fused loader.
        Returns [B*G, ...] in `out_format` (G images per sample, or only the first when first_only); `out` makes the
        loader write into a caller-owned tensor of that shape (the executor's fixed-address step buffer)."""
        if "jpeg_blob" in batch:
            B, G, S = batch["codes"].shape[0], batch["codes"].shape[1], batch["size"]
            select = None
            if first_only:      # decode only the first image's six files of every sample
                select = (torch.arange(B, device=dev)[:, None] * (G * 6) + torch.arange(6, device=dev)[None]).flatten()
            planes = ops.jpeg_decode_gray(batch["jpeg_blob"].to(dev, non_blocking=True),
                                          batch["jpeg_offsets"].to(dev, non_blocking=True), (S, S), select=select)
            planes = planes.view(B, 1 if first_only else G, 6, S, S)
        else:
            planes = batch["planes"].to(dev, non_blocking=True)                  # [B,G,6,H,W] u8
        B, G = planes.shape[:2]
        codes, crops = batch["codes"].to(dev), batch["crops"].to(dev)
        mats = batch["mats"].to(dev) if "mats" in batch else None
        if first_only:
            planes, codes, crops, G = planes[:, :1], codes[:, :1], crops[:, :1], 1
            mats = mats[:, :1] if mats is not None else None
        planes = planes.reshape(B * G, *planes.shape[2:]).contiguous()
        exp = batch["exp"].to(dev).to(torch.int32).repeat_interleave(G)
        norm_m, norm_d = self._norm(dev)
        hw = batch["out"]
        if mats is not None:
            return ops.load_norm_affine(planes, torch.arange(B * G, dtype=torch.int32, device=dev), exp,
                                        codes.reshape(-1).contiguous(), mats.reshape(-1, 2, 3).contiguous(),
                                        crops.reshape(-1, 2).contiguous(), norm_m, norm_d, (hw, hw), out_format, out=out)
        return ops.load_norm_aug(planes, torch.arange(B * G, dtype=torch.int32, device=dev), exp,
                                 codes.reshape(-1).contiguous(), crops.reshape(-1, 2).contiguous(), norm_m, norm_d,
                                 (hw, hw), out_format, out=out)

    # ------------------------------------------------------------ reference-compatible item (dataloader.py:148-209)
    def __getitem__(self, index):
        item = self.raw_item(index)
        dev = torch.device(self.device)
        batch = collate_raw([item])
        x = self.device_batch(batch, dev, out_format=ops.OUT_F32_NCHW)           # [G,6,h,w] float32, bit-exact fp32 mode
        return x.cpu(), item["label"]

    def __len__(self):
        return self.len


def collate_raw(items):
    out = _collate_common(items)
    if "mats" in items[0]:
        out["mats"] = torch.stack([it["mats"] for it in items])
    return out


def _collate_common(items):
    if "jpeg" in items[0]:
        blob, offsets = ops.pack_jpeg_buffers([b for it in items for b in it["jpeg"]])
        head = {"jpeg_blob": blob, "jpeg_offsets": offsets, "size": items[0]["size"]}
    else:
        head = {"planes": torch.stack([it["planes"] for it in items])}
    return {**head,
            "codes": torch.stack([it["codes"] for it in items]),
            "crops": torch.stack([it["crops"] for it in items]),
            "exp": torch.tensor([it["exp"] for it in items], dtype=torch.int32),
            "out": items[0]["out"],
            "label": [it["label"] for it in items]}


class RawView(torch.utils.data.Dataset):
    """What train()/test() hand to torch's DataLoader: worker processes only decode.  controls=False leaves the
    control wells out of the items (see ImagesDS.raw_item)."""

    def __init__(self, ds, controls=True):
        self.ds = ds
        self.controls = controls

    def __len__(self):
        return len(self.ds)

    def __getitem__(self, i):
        return self.ds.raw_item(i, controls=self.controls)


def train_test_split(df, random_state):
    """Split by experiment within each cell type — the reference's helper of the same name (dataloader.py:215-239).
    Host-side pandas glue, outside the hot path; kept only because main.py:13-15 imports it from this module
    unconditionally (an unchanged main.py would not start without it) and :103 calls it when
    HYPERPARAMS['train_split_by_experiment'] is set: a third of each cell
    type's experiments (column 'exp'), drawn with random.shuffle under the given seed, goes to validation; both frames
    are then shuffled with the same seed."""
    import pandas as pd
    random.seed(random_state)
    train_parts, val_parts = [], []
    for celltype in df['celltype'].unique():
        part = df[df['celltype'] == celltype]
        exps = part['exp'].unique()
        n_val = len(exps) // 3
        random.shuffle(exps)
        in_val = part['exp'].isin(list(exps[:n_val]))
        train_parts.append(part[~in_val])
        val_parts.append(part[in_val])
    shuffle = lambda parts: pd.concat(parts).sample(frac=1, random_state=random_state).reset_index(drop=True)
    return shuffle(train_parts), shuffle(val_parts)
