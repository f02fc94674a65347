"""CPU restatement of the reference's motion-vector / partition rasteriser (ORACLE, test infrastructure).

Only ``tests/`` may import this file.  It follows, statement by statement, the per-record loop of
``LoadImageFromFileList_ipb.__call__`` (mmedit/datasets/pipelines/loading_ipb.py:328-369, record layout
:339, p_offset bookkeeping :369) with ``load_partition=True, drconv=True`` as in the shipped configs
(configs/HR_davis_LR_128x128_IPB.py:63-65), followed by the ``RescaleToZeroOne`` division of the partition
maps by 255 (mmedit/datasets/pipelines/normalization.py:93-99) and the HWC->CHW move of ``FramesToTensor``
(formating.py:101-138).

Parity status: PINNED -- ``tests/golden/make_golden_raster.py`` executes the reference's own loop (its
source lines are read from /root/reference at generation time, not copied) on seeded records and commits
the results; ``tests/test_sideinfo.py`` holds this restatement to bit equality against them.

A record is ``[direction, w, h, src_x, src_y, dst_x, dst_y, motion_x, motion_y, scale]`` (float32).
"""
import numpy as np

PARTITION_CHANNEL = {256: 0, 128: 1, 64: 2}      # loading_ipb.py:334 ({'256':0,'128':1,'64':2}[str(w*h)])


def rasterize_clip(records, slice_types, h_img, w_img):
    """records: list (one per frame) of (R_f, 10) arrays; slice_types: list of 'I' | 'P' | 'B'.

    Returns (mvs (T,4,H,W) float32, partitions (T,3,H,W) float32 already divided by 255).
    Raises like the reference: KeyError for a block area outside {256,128,64}; UnboundLocalError when a
    reverse (P-frame) record appears before ``p_offset`` exists (first frame), IndexError when it points
    before the clip.
    """
    mvs, partitions = [], []
    p_offset = None                                        # unbound in the reference until :369 runs
    for rec_f, st in zip(records, slice_types):
        is_b = (st == "B")                                 # :316
        mv_npy = np.asarray(rec_f, dtype=np.float32).reshape(-1, 10)   # :329 .astype(np.float32)
        mv = np.zeros((h_img, w_img, 4)).astype(np.float32)            # :331
        partition = np.zeros((h_img, w_img, 3)).astype(np.float32)     # :334
        for idx in range(mv_npy.shape[0]):
            direction, w, h, x_w, y_w, x, y, motion_x, motion_y, scale = mv_npy[idx]      # :339
            x, y, w, h, x_w, y_w = int(x), int(y), int(w), int(h), int(x_w), int(y_w)     # :340
            motion_x = motion_x / scale                                                   # :341 (float32)
            motion_y = motion_y / scale                                                   # :342
            if direction < 0:                                                             # :343-346
                mv[y - h // 2:y + h // 2, x - w // 2:x + w // 2, 0] = motion_x
                mv[y - h // 2:y + h // 2, x - w // 2:x + w // 2, 1] = motion_y
            elif direction > 0 and is_b:                                                  # :347-350
                mv[y - h // 2:y + h // 2, x - w // 2:x + w // 2, 2] = motion_x
                mv[y - h // 2:y + h // 2, x - w // 2:x + w // 2, 3] = motion_y
            elif direction > 0 and (not is_b):                                            # :351-355
                if p_offset is None:
                    raise UnboundLocalError("p_offset referenced before assignment")
                tgt = mvs[-p_offset]
                tgt[y_w - h // 2:y_w + h // 2, x_w - w // 2:x_w + w // 2, 2] = -motion_x
                tgt[y_w - h // 2:y_w + h // 2, x_w - w // 2:x_w + w // 2, 3] = -motion_y
            # direction == 0: the reference's `assert TypeError(...)` is a no-op           # :356-357
            partition[y - h // 2:y + h // 2, x - w // 2:x + w // 2, PARTITION_CHANNEL[w * h]] = 1   # :361
        partitions.append(partition)                                                      # :366
        mvs.append(mv)                                                                    # :368
        p_offset = p_offset + 1 if (is_b and p_offset is not None) else 1                 # :369
        # (the reference's first frame is never 'B' in practice; p_offset+1 on an unbound name would
        #  raise there too -- treated as 1 here only for clips that start with B and never use it)
    mvs = np.stack(mvs, 0).transpose(0, 3, 1, 2).copy()                                   # FramesToTensor
    partitions = (np.stack(partitions, 0).astype(np.float32) / 255).transpose(0, 3, 1, 2).copy()
    return mvs.astype(np.float32), partitions.astype(np.float32)
