"""Converts the reference's ASCII force-kernel tables (golden *inputs*) into small
binary fixtures.  Run in the build container (needs /root/reference); the .npy files are committed.

  wfxyzf.3.ascii  (CUBE/kernels, read by CUBE/main/kernel_f.f90:14-23)  -> fk_table.npy  [k][j][i][dim] (16,16,16,3) f32
  wfxyzc.2.ascii  (CUBE/kernels, read by CUBE/main/kernel_c.f90:30-38)  -> ck_table.npy  [k][j][i][dim] (4,4,4,3)   f32
"""
import hashlib
import os
import numpy as np

REF = "/root/reference/CUBE/kernels"
HERE = os.path.dirname(os.path.abspath(__file__))
MD5 = {"wfxyzf.3.ascii": "9b2c7e4219615cf3efec762c0a23e807", "wfxyzc.2.ascii": "f3f8ecf8dd0766093d15a67c772e65e9"}


def load(name, n):
    path = os.path.join(REF, name)
    assert hashlib.md5(open(path, "rb").read()).hexdigest() == MD5[name], name
    a = np.loadtxt(path)
    assert a.shape == (n ** 3, 6)
    idx = a[:, :3].astype(int).reshape(n, n, n, 3)
    k, j, i = np.meshgrid(np.arange(1, n + 1), np.arange(1, n + 1), np.arange(1, n + 1), indexing="ij")
    assert (idx[..., 0] == i).all() and (idx[..., 1] == j).all() and (idx[..., 2] == k).all()  # kernel_f.f90:19
    return a[:, 3:].astype(np.float32).reshape(n, n, n, 3)


if __name__ == "__main__":
    np.save(os.path.join(HERE, "fk_table.npy"), load("wfxyzf.3.ascii", 16))
    np.save(os.path.join(HERE, "ck_table.npy"), load("wfxyzc.2.ascii", 4))
    print("ok")
