"""oracle/prepare_ref.py -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

numpy restatement of the reference's volume -> point cloud preparation (the step before the hot path), loop for loop where the
reference loops, with the random draws INJECTED (the reference uses Python's ``random.sample`` / ``DP.shuffle_idx``):

  intensity_normalize_whole     utils/dataPreparePancreas.py:32-46   (mean / std over the whole volume)
  intensity_normalize_nonzero   utils/dataPrepareBraTS.py:32-49      (mean / std over v > 0, zeros stay 0)
  pancreas_cloud                utils/dataPreparePancreas.py:132-169 (sampling_convert_pc2ply, one "loop")
  brats_full_cloud              utils/dataPrepareBraTS.py:75-94      (convert_pc2ply up to the full cloud)
  brats_sample                  runBraTS.py:104-119                  (tumour + drawn non-tumour points, shuffled)
"""
import numpy as np


def intensity_normalize_whole(volume):
    pixels = volume
    return (volume - pixels.mean()) / pixels.std()


def intensity_normalize_nonzero(volume):
    pixels = volume[volume > 0]
    out = (volume - pixels.mean()) / pixels.std()
    out[volume == 0] = 0.0
    return out


def pancreas_cloud(img, label, n_point, background_choice):
    """``background_choice``: positions in ``none_tumor`` (the reference draws them with random.sample)."""
    x_axis, y_axis, z_axis = img.shape
    data_list = [[x, y, z, img[x][y][z], label[x][y][z]] for x in range(x_axis) for y in range(y_axis) for z in range(z_axis)]
    pc_label = np.array(data_list)
    xyz_min = np.array([x_axis, y_axis, z_axis]).astype(np.float32)
    xyz = pc_label[:, :3].astype(np.uint16)
    colors = pc_label[:, 3:4].astype(np.float32)
    labels = pc_label[:, 4].astype(np.uint8)
    none_tumor = list(np.where(labels == 0)[0])
    tumor = list(np.where(labels > 0)[0])
    assert len(background_choice) == n_point - len(tumor)
    queried_idx = np.array(tumor + [none_tumor[i] for i in background_choice])
    sampling_xyz = xyz[queried_idx].astype(np.uint16)                 # saved as <ID>_xyz_origin_loop_<i>.npy
    return dict(xyz_origin=sampling_xyz, xyz=sampling_xyz.astype(np.float32) / xyz_min, value=colors[queried_idx],
                labels=labels[queried_idx])


def brats_full_cloud(volume):
    """``volume [5,X,Y,Z]`` float64: 4 z-scored modalities + label."""
    channel, x_axis, y_axis, z_axis = volume.shape
    data_list = [[x, y, z, volume[0][x][y][z], volume[1][x][y][z], volume[2][x][y][z], volume[3][x][y][z], volume[4][x][y][z]]
                 for x in range(x_axis) for y in range(y_axis) for z in range(z_axis)
                 if (volume[0][x][y][z] != 0 or volume[1][x][y][z] != 0 or volume[2][x][y][z] != 0 or volume[3][x][y][z] != 0)]
    pc_data = np.array(data_list)
    xyz_origin = pc_data[:, :3].astype(int)                            # saved as <ID>_xyz_origin.npy
    xyz_min = np.array([x_axis, y_axis, z_axis])
    pc_data[:, 0:3] /= xyz_min
    return dict(xyz_origin=xyz_origin, xyz=pc_data[:, :3].astype(np.float32), colors=pc_data[:, 3:7].astype(np.float32),
                labels=pc_data[:, 7].astype(np.uint8))


def brats_sample(full, num_points, background_choice, shuffle_perm):
    """runBraTS.py:104-119 with the random.sample positions and the shuffle permutation injected."""
    all_label = full["labels"]
    none_tumor = list(np.where(all_label == 0)[0])
    tumor = list(np.where(all_label > 0)[0])
    queried_idx = np.array(tumor + [none_tumor[i] for i in background_choice])
    assert len(queried_idx) == num_points
    queried_idx = queried_idx[shuffle_perm]
    return dict(xyz=full["xyz"][queried_idx].astype(np.float32), colors=full["colors"][queried_idx].astype(np.float32),
                labels=full["labels"][queried_idx], point_idx=queried_idx.astype(np.int32))
