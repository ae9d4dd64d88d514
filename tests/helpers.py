"""Shared test helpers: the oracle version of the binarization_soma chain."""
import numpy as np

import oracle


def oracle_chain(case, nms_thresh=0.23, keep_largest_cc=True):
    """binarization_soma.py:57-104 with oracle pieces (keep_largest_cc: the :97-99 filter, scipy restatement).
    Returns dict(seg, order, status{inst: code}, b_max{inst: b}, survive [by rank])."""
    dets, boxes, prm, off, vol = case["dets"], case["boxes"], case["prm"], case["crop_off"], case["volume"]
    keep = oracle.nms_3d(dets, nms_thresh)
    order = keep[oracle.argsort_desc(dets[keep, 6])] if len(keep) else keep
    seg = np.zeros(vol.shape, np.uint16)
    status, bmax = {}, {}
    pm, pid, pbox, prank = [], [], [], []
    for rank, i in enumerate(order):
        b = boxes[i]
        img = vol[b[2]:b[5] + 1, b[1]:b[4] + 1, b[0]:b[3] + 1]
        p = prm[off[i]:off[i + 1]].reshape(img.shape)
        if p.max() == 0:
            status[int(i)] = 3
            continue
        i16, p16 = oracle.soma_normalise(img, p)
        try:
            m, _, bm = oracle.otsu_py_2d_fast(i16, p16)
        except UnboundLocalError:
            status[int(i)] = 1
            continue
        bmax[int(i)] = bm
        if keep_largest_cc:
            try:
                m = (oracle.largest_cc(m) * 255).astype(np.uint8)
            except IndexError:                              # no foreground: the reference raises here
                status[int(i)] = 5
                continue
        status[int(i)] = 0
        pm.append(m); pid.append(rank + 1); pbox.append(b); prank.append(rank)
    surv = np.zeros(len(order), bool)
    if pm:
        s = oracle.paste_labels(seg, np.asarray(pbox), np.asarray(pid, np.uint16), pm)
        surv[np.asarray(prank)] = s
    return dict(seg=seg, order=np.asarray(order, np.int64), status=status, b_max=bmax, survive=surv)
