"""Cross-tile de-duplication of detections — the reference's full-frame inference mode.

Mirrors ``src/visualize.py``: a 704x520 frame is a 7x7 grid of mini-tiles, inference runs on 25 overlapping
3x3-mini-tile windows, and a detection of tile t is kept when more than ``mask_threshold`` of its mask area lies in
the mini-tiles tile t is responsible for (its centre, or grid-border mini-tiles not claimed by an earlier tile)
(``filter_detections_by_border_mini_tiles``, visualize.py:174-257; ``calculate_mask_area_in_region``, :106-130).

The reference moves every prediction to the CPU and loops over detections x regions in numpy; here the masks stay
on the GPU and one kernel per tile returns exact integer pixel counts (``lcr_mask_region_counts_u8``).  The float64
fractions, their summation order and the threshold are evaluated on the host exactly as the reference does, so the
kept set, the boxes and the reported ``area_fraction`` are identical.
"""
from __future__ import annotations

import torch

from . import ops
from ._lib import LcrError

IMG_WIDTH = 704
IMG_HEIGHT = 520
N_MINI_COLS = 7
N_MINI_ROWS = 7
TILE_SIZE = 3
N_TILES_COL = N_MINI_COLS - TILE_SIZE + 1
N_TILES_ROW = N_MINI_ROWS - TILE_SIZE + 1
TOTAL_TILES = N_TILES_COL * N_TILES_ROW


def get_tile_position_in_grid(tile_num):
    """(col_start, row_start) of tile `tile_num` in mini-tile units (visualize.py:100-103)."""
    return tile_num % N_TILES_COL, tile_num // N_TILES_COL


def get_valid_mini_tiles_for_tile(tile_num):
    """Centre mini-tile of the 3x3 window plus the mini-tiles on the grid border (visualize.py:151-171)."""
    col_start, row_start = get_tile_position_in_grid(tile_num)
    valid = []
    for local_row in range(TILE_SIZE):
        for local_col in range(TILE_SIZE):
            mini_col, mini_row = col_start + local_col, row_start + local_row
            is_center = local_col == 1 and local_row == 1
            on_border = mini_col == 0 or mini_col == N_MINI_COLS - 1 or mini_row == 0 or mini_row == N_MINI_ROWS - 1
            if is_center or on_border:
                valid.append((mini_col, mini_row))
    return valid


def filter_detections_by_border_mini_tiles(results, score_threshold=0.5, mask_threshold=0.4, masks_to_host=False):
    """Same contract as visualize.py:174-257.  results: list of {'tile_num', 'prediction': {'boxes','scores','masks'}}
    with CUDA tensors (masks uint8 [N,h,w] in {0,255}, or float [N,1,h,w] probabilities as torchvision returns).
    Returns the reference's list of dicts; 'mask' is a bool CUDA tensor (numpy when masks_to_host)."""
    mtw, mth = IMG_WIDTH // N_MINI_COLS, IMG_HEIGHT // N_MINI_ROWS
    out, processed = [], set()
    for result in sorted(results, key=lambda x: x["tile_num"]):
        tile_num, pred = result["tile_num"], result["prediction"]
        col_start, row_start = get_tile_position_in_grid(tile_num)
        off_x, off_y = col_start * mtw, row_start * mth
        new_tiles = [mt for mt in get_valid_mini_tiles_for_tile(tile_num) if mt not in processed]
        if not new_tiles:
            continue
        if not pred["masks"].is_cuda:
            raise LcrError("stitch: predictions must stay on the GPU (there is no CPU fallback)")
        keep = pred["scores"] > score_threshold
        boxes, scores, masks = pred["boxes"][keep], pred["scores"][keep], pred["masks"][keep]
        if masks.dim() == 4:
            masks = masks[:, 0]
        mbool = masks > 0.5 if masks.dtype != torch.uint8 else masks > 0
        n = int(boxes.shape[0])
        if n:
            h, w = int(mbool.shape[-2]), int(mbool.shape[-1])
            rects, live = [], []
            for mc, mr in new_tiles:                       # region in tile-local coordinates, clipped (visualize.py:109-124)
                x0, y0 = max(0, mc * mtw - off_x), max(0, mr * mth - off_y)
                x1, y1 = min(w, mc * mtw + mtw - off_x), min(h, mr * mth + mth - off_y)
                live.append(x0 < x1 and y0 < y1)
                rects.append([x0, y0, x1, y1] if live[-1] else [0, 0, 0, 0])
            R = len(rects)
            rt = torch.tensor(rects, dtype=torch.int32, device=boxes.device).repeat(n, 1)
            ro = torch.arange(0, (n + 1) * R, R, dtype=torch.int32, device=boxes.device)
            total, inreg = ops.mask_region_counts(mbool.to(torch.uint8), rt, ro)
            total, inreg = total.tolist(), inreg.reshape(n, R).tolist()
            b_host, s_host = boxes.tolist(), scores.tolist()
            for i in range(n):
                frac = 0.0
                for r in range(R):                         # float(pixels_in_region / total_pixels), summed in region order
                    frac += (inreg[i][r] / total[i]) if (live[r] and total[i] > 0) else 0.0
                if frac > mask_threshold:
                    bx = b_host[i]
                    out.append({
                        "box": [bx[0] + off_x, bx[1] + off_y, bx[2] + off_x, bx[3] + off_y],
                        "mask": mbool[i].cpu().numpy() if masks_to_host else mbool[i],
                        "score": float(s_host[i]),
                        "tile_num": tile_num,
                        "offset": (off_x, off_y),
                        "area_fraction": frac,
                        "mini_tile": new_tiles,
                    })
        processed.update(new_tiles)
    return out
