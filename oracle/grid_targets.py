"""TEST INFRASTRUCTURE ONLY -- numpy restatement of GridLossComputation.prepare_target
(pet/rcnn/modeling/grid_cascade_rcnn/loss.py:178-258): fp32 arithmetic in the reference's order, int() truncation."""
import numpy as np

F = np.float32


def prepare_target(pos_bboxes, pos_gt_bboxes, mapping_ratio, sub_regions, pos_radius=1, grid_points=9, map_size=56,
                   target_refine=False):
    b, g = np.asarray(pos_bboxes, F).reshape(-1, 4), np.asarray(pos_gt_bboxes, F).reshape(-1, 4)
    ratio = F(mapping_ratio)
    x1 = b[:, 0] - ratio * ((b[:, 2] - b[:, 0]) / F(2))
    y1 = b[:, 1] - ratio * ((b[:, 3] - b[:, 1]) / F(2))
    x2 = b[:, 2] + ratio * ((b[:, 2] - b[:, 0]) / F(2))
    y2 = b[:, 3] + ratio * ((b[:, 3] - b[:, 1]) / F(2))
    ws, hs = x2 - x1, y2 - y1
    R = b.shape[0]
    gs = int(np.sqrt(grid_points))
    targets = np.zeros((R, grid_points, map_size, map_size), F)
    r2 = pos_radius ** 2
    for i in range(R):
        if ws[i] <= gs or hs[i] <= gs:
            continue
        for j in range(grid_points):
            fx, fy = 1 - (j // gs) / (gs - 1), 1 - (j % gs) / (gs - 1)          # python floats, cast when multiplied
            gx = F(fx) * g[i, 0] + F(1 - fx) * g[i, 2]
            gy = F(fy) * g[i, 1] + F(1 - fy) * g[i, 3]
            cx = int((gx - x1[i]) / ws[i] * F(map_size))
            cy = int((gy - y1[i]) / hs[i] * F(map_size))
            for x in range(cx - pos_radius, cx + pos_radius + 1):
                for y in range(cy - pos_radius, cy + pos_radius + 1):
                    if 0 <= x < map_size and 0 <= y < map_size and (x - cx) ** 2 + (y - cy) ** 2 <= r2:
                        targets[i, j, y, x] = 1
            if target_refine and (cx < 0 or cx >= map_size or cy < 0 or cy >= map_size):
                targets[i, j, min(max(cy, 0), map_size - 1), min(max(cx, 0), map_size - 1)] = 1
    return np.concatenate([targets[:, [j], s[1]:s[3], s[0]:s[2]] for j, s in enumerate(sub_regions)], axis=1)
