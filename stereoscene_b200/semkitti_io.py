"""SemanticKITTI prediction writer -- the last step after the path when predictions are saved instead of scored
(reference: ``save_output_semantic_kitti``, projects/mmdet3d_plugin/occupancy/apis/test.py:49-64, with
``get_inv_map`` from projects/mmdet3d_plugin/utils/semkitti_io.py:99-111 reading ``learning_map_inv`` of
semantickitti.yaml:144-164).  Host-side I/O: the uint8 label volume of the trilinear + argmax kernel is remapped to
the lidarseg ids and written as little-endian uint16, byte-identical to the reference's file.
"""
from __future__ import annotations

import os

import numpy as np

# semantickitti.yaml:144-164 (learning_map_inv): training class index -> SemanticKITTI lidarseg id
LEARNING_MAP_INV = {0: 0, 1: 10, 2: 11, 3: 15, 4: 18, 5: 20, 6: 30, 7: 31, 8: 32, 9: 40, 10: 44, 11: 48, 12: 49, 13: 50,
                    14: 51, 15: 70, 16: 71, 17: 72, 18: 80, 19: 81}


def get_inv_map() -> np.ndarray:
    """int32[20] lookup table, same values as the reference's ``get_inv_map`` (which needs the yaml in the CWD)."""
    inv_map = np.zeros(20, dtype=np.int32)
    inv_map[list(LEARNING_MAP_INV.keys())] = list(LEARNING_MAP_INV.values())
    return inv_map


def save_output_semantic_kitti(output_voxels, save_path: str, sequence_id: str, frame_id: str, verbose: bool = False) -> str:
    """Write ``{save_path}/sequences/{sequence_id}/predictions/{frame_id}.label``.

    ``output_voxels``: either the class logits ``[C, X, Y, Z]`` (the reference's argument: argmax over dim 0 is taken
    here, as in test.py:52) or an integer label volume ``[X, Y, Z]`` (e.g. ``out["labels"][b]`` from
    ``forward_features(..., want_labels=True)``, which already is that argmax).  Torch tensors (any device) or numpy."""
    if hasattr(output_voxels, "detach"):
        t = output_voxels.detach()
        if t.is_floating_point():
            t = t.argmax(dim=0)
        labels = t.cpu().numpy()
    else:
        labels = np.asarray(output_voxels)
        if np.issubdtype(labels.dtype, np.floating):
            labels = labels.argmax(axis=0)
    labels = labels.reshape(-1).astype(np.int64)
    if labels.size and (labels.min() < 0 or labels.max() >= 20):
        raise ValueError("save_output_semantic_kitti: labels outside [0, 20)")
    out = get_inv_map()[labels].astype(np.uint16)
    folder = "{}/sequences/{}/predictions".format(save_path, sequence_id)
    os.makedirs(folder, exist_ok=True)
    path = os.path.join(folder, "{}.label".format(frame_id))
    with open(path, "wb") as f:
        out.tofile(f)
    if verbose:
        print("\n save to {}".format(path))
    return path
