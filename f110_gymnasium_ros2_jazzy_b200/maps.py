"""Map loading: yaml + image -> fp64 distance transform, on the host, as ScanSimulator2D.set_map does
(laser_models.py:383-427).  The EDT stays on scipy; only the upload is new."""
import os

import numpy as np


def load_map(map_path, map_ext):
    """-> (dt fp64 [H, W] with row 0 = bottom image row, resolution, [ox, oy, oyaw])."""
    import yaml
    from PIL import Image
    from scipy.ndimage import distance_transform_edt
    map_img_path = os.path.splitext(map_path)[0] + map_ext
    img = np.array(Image.open(map_img_path).transpose(Image.FLIP_TOP_BOTTOM)).astype(np.float64)
    img[img <= 128.] = 0.
    img[img > 128.] = 255.
    with open(map_path, 'r') as stream:
        meta = yaml.safe_load(stream)
    resolution = meta['resolution']
    origin = meta['origin']
    dt = resolution * distance_transform_edt(img)
    return dt, float(resolution), [float(origin[0]), float(origin[1]), float(origin[2])]


def load_map_image(map_path, map_ext):
    """-> (free mask uint8 [H, W] with row 0 = bottom image row, resolution, [ox, oy, oyaw]): the input of the EDT, binarised
    exactly as laser_models.py:398-404 does (pixel <= 128 obstacle, > 128 free)."""
    import yaml
    from PIL import Image
    map_img_path = os.path.splitext(map_path)[0] + map_ext
    img = np.array(Image.open(map_img_path).transpose(Image.FLIP_TOP_BOTTOM)).astype(np.float64)
    with open(map_path, 'r') as stream:
        meta = yaml.safe_load(stream)
    origin = meta['origin']
    return (img > 128.).astype(np.uint8), float(meta['resolution']), [float(origin[0]), float(origin[1]), float(origin[2])]


def map_bounds(map_path, map_dir=None):
    """World-frame bounds used for the observation space (f110_env.py:224-232)."""
    import yaml
    from PIL import Image
    meta = yaml.safe_load(open(map_path))
    res = meta['resolution']
    x0, y0, _ = meta.get('origin', (0.0, 0.0, 0.0))
    base = map_dir if map_dir is not None else os.path.dirname(map_path) + '/'
    width, height = Image.open(base + meta['image']).size
    return x0, x0 + width * res, y0, y0 + height * res
