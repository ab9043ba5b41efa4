"""oracle/improcess_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Float64 restatement of the image-domain ("Gabor") detector of DAS4Whales: the functions of
/root/reference/src/das4whales/improcess.py that scripts/main_gabordetect.py:78-169 calls, plus that script section itself.
Like the reference it sits on OpenCV (cv2.getGaborKernel, cv2.filter2D) and torchvision (Resize) -- the reference's own
third-party dependencies, both importable in this image -- so it is the reference arithmetic, not a re-derivation.
Pinned against the unmodified reference module by oracle/make_golden.py (tests/golden/gabor.npz).
Only tests/, smoke() and bench.py's CPU legs may import this.
"""
import numpy as np
import scipy.signal as sps


def scale_pixels(img):
    """improcess.py:23-41"""
    return (img - img.min()) / (img.max() - img.min())


def trace2image(trace):
    """improcess.py:44-63"""
    image = np.abs(sps.hilbert(trace, axis=1)) / np.std(trace, axis=1, keepdims=True)
    return scale_pixels(image) * 255


def angle_fromspeed(c0, fs, dx, selected_channels):
    """improcess.py:66-95 (without the prints)"""
    return np.arctan(c0 / (fs * dx * selected_channels[2])) * 180 / np.pi


def gabor_filt_design(theta_c0):
    """improcess.py:98-140"""
    import cv2
    theta = np.pi / 2 + np.deg2rad(theta_c0)
    up = cv2.getGaborKernel((100, 100), 4, theta, 20, 0.15, 0, ktype=cv2.CV_64F)
    return up, np.flipud(up)


def binning(image, ft, fx):
    """improcess.py:395-421 -- torchvision ToTensor + Resize (bilinear, antialias)"""
    import torchvision.transforms as transforms
    t = transforms.ToTensor()(image)
    t = transforms.Resize((int(image.shape[0] * fx), int(image.shape[1] * ft)))(t)
    return t.numpy()[0]


def filter2D(img, kernel):
    """cv2.filter2D(img, cv2.CV_64F, kernel) -- scripts/main_gabordetect.py:109,135"""
    import cv2
    return cv2.filter2D(np.ascontiguousarray(img, dtype=np.float64), cv2.CV_64F, np.ascontiguousarray(kernel))


def apply_smooth_mask(array, mask):
    """improcess.py:424-454: the product uses the raw mask (:452)"""
    return array * mask


def gabor_detect(trf_fk, fs, dx, selected_channels, c0=1500., bin_factor=10, threshold=9100., threshold2=150.):
    """scripts/main_gabordetect.py:78-169"""
    image = trace2image(trf_fk)
    theta_c0 = angle_fromspeed(c0, fs, dx, selected_channels)
    imagebin = binning(image, 1 / bin_factor, 1 / bin_factor)
    up, down = gabor_filt_design(theta_c0)
    fimage = filter2D(imagebin, up) + filter2D(imagebin, down)
    binary = fimage > threshold
    m2 = filter2D(binary.astype(float), up) + filter2D(binary.astype(float), down)
    mask = m2 > threshold2
    mask_sparse = binning(mask, bin_factor, bin_factor)
    masked = apply_smooth_mask(trf_fk, mask_sparse)
    return masked, {"image": image, "imagebin": imagebin, "fimage": fimage, "m2": m2, "mask": mask, "mask_sparse": mask_sparse}
