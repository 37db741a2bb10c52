"""Generates tests/golden/*.npz from the reference's bundled logs (run in the authoring container,
where /root/reference exists; the GPU box only sees the committed .npz files).

    python tests/golden/make_golden.py

Contents (all float64, parsed from the 6-significant-digit text logs with numpy.loadtxt):
  water_corners [1062,18]  t id + 16 undistorted normalised stereo corner coords  (vision.cpp:111-119)
  water_image   [1062,9]   t id p(3) q(wxyz): logged output of RefractionTriangulation+ComputeMarkerPose
  land_corners  [1257,14]  t id + 4 triangulated 3-D corners                    (vision.cpp:120-124)
  land_image    [1257,9]   logged output of ComputeMarkerPose
  land_imu / water_imu     [n,7] t accel(3) gyro(3) raw IMU log                  (filter.cpp:31-32)
  land_fusion / water_fusion [n,17] logged filter output of an OLDER reference revision: shape-only pin
Source: /root/reference/matlab/dataset/{landdata/dataset-02,waterdata/dataset-06}.

Also writes log_heads.json: the first 4 text lines of every log verbatim (without the line ends), the format pin of
fbus_ekf_b200/logio.py (`ofstream <<` at default precision) where /root/reference does not exist.
"""
import json
import os
import numpy as np

SRC = "/root/reference/matlab/dataset"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    out = {}
    for key, sub in (("land", "landdata/dataset-02"), ("water", "waterdata/dataset-06")):
        for name in ("corners", "image", "imu", "fusion"):
            out[f"{key}_{name}"] = np.loadtxt(os.path.join(SRC, sub, name + ".txt"))
    np.savez_compressed(os.path.join(HERE, "fbus_logs.npz"), **out)
    heads = {}
    for key, sub in (("land", "landdata/dataset-02"), ("water", "waterdata/dataset-06")):
        for name in ("corners", "image", "imu", "fusion"):
            with open(os.path.join(SRC, sub, name + ".txt"), newline="") as fh:
                heads[f"{key}_{name}"] = [fh.readline().rstrip("\r\n") for _ in range(4)]
    json.dump(heads, open(os.path.join(HERE, "log_heads.json"), "w"), indent=1)
    for k, v in out.items():
        print(k, v.shape)


if __name__ == "__main__":
    main()
