"""Prints the cv::cartToPolar table behind the HOG restatement (oracle/oracle_learn.cpp kMagBits / kAngBits and the
same table in agile_grasp_b200/csrc/hog_svm.cu): magnitude and angle of the 9 gradient cases (dx, dy) in
{-s, 0, +s}^2, s = sqrt(255.f), of a binary 0/255 image after gamma correction, as computed by cv2 itself
(OpenCV uses a polynomial atan, so the diagonals are not exact multiples of pi/4).  Run in the build container;
tests/test_oracle_hog_svm.py::test_cart_to_polar_table_is_what_cv2_computes pins the committed constants."""
import cv2
import numpy as np

s = np.float32(np.sqrt(np.float32(255.0)))
sx = np.array([[-1, 0, 1]] * 3, np.float32)
sy = np.array([[-1] * 3, [0] * 3, [1] * 3], np.float32)
mag, ang = cv2.cartToPolar(sx * s, sy * s)
fmt = lambda a: ", ".join(hex(int(v)) for v in a.ravel().view(np.uint32))
print("// index = (sy+1)*3 + (sx+1)")
print("kMagBits = {" + fmt(mag) + "};")
print("kAngBits = {" + fmt(ang) + "};")
