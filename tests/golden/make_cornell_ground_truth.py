"""Generates tests/golden/cornell_box_tungsten_256.npy: the reference's ground-truth render
/root/reference/scenes/cornell-box/TungstenRender.exr (1024x1024 linear RGB) box-downsampled to
256x256, float16.  Run in the build container (the GPU box has no /root/reference)."""
import os
os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
import cv2
import numpy as np

src = "/root/reference/scenes/cornell-box/TungstenRender.exr"
im = cv2.imread(src, cv2.IMREAD_UNCHANGED)[..., ::-1].astype(np.float64)      # BGR -> RGB
h, w, _ = im.shape
f = h // 256
small = im.reshape(256, f, 256, f, 3).mean(axis=(1, 3))
np.save(os.path.join(os.path.dirname(os.path.abspath(__file__)), "cornell_box_tungsten_256.npy"), small.astype(np.float16))
print(small.shape, small.mean(axis=(0, 1)))
