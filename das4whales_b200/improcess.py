"""das4whales_b200.improcess -- the image-domain ("Gabor") detector of DAS4Whales on the GPU.

Drop-in for the functions of `das4whales.improcess` that scripts/main_gabordetect.py calls
(/root/reference/src/das4whales/improcess.py): scale_pixels (:23-41), trace2image (:44-63), angle_fromspeed (:66-95),
gabor_filt_design (:98-140), binning (:395-421), apply_smooth_mask (:424-454); plus `filter2D`, the GPU stand-in for the
`cv2.filter2D(img, cv2.CV_64F, kernel)` calls the script makes (:109, :135), and `gabor_detect`, which chains the whole
script section :78-169 on the device.  Everything numeric runs in libd4w.so (csrc/image_kernels.cuh); OpenCV / torchvision
are not used.

ndarray in -> float64 ndarray out (like the reference); CUDA tensor in -> float32 CUDA tensor out.
"""
import numpy as np

from . import _lib
from . import rows as _rows
from .dsp import _is_tensor, _to_device, _to_host64


def _torch():
    return _rows._torch()


def _dev2d(a):
    """ndarray / tensor (any real or bool dtype) -> contiguous float32 CUDA tensor [h, w]"""
    torch = _torch()
    if _is_tensor(a):
        t = a if a.is_cuda else a.cuda()
        return t.to(torch.float32).contiguous()
    arr = np.asarray(a)
    if arr.dtype == np.bool_:
        arr = arr.astype(np.float32)
    return _to_device(arr)


def scale_pixels(img):
    """(img - min) / (max - min)  (reference: improcess.py:23-41)"""
    return _scale(img, 1.0)


def _scale(img, mul):
    torch = _torch()
    x = _dev2d(img)
    y = torch.empty_like(x)
    ws = torch.empty(2, dtype=torch.int32, device=x.device)
    with torch.cuda.device(x.device.index):
        _lib.check(_lib.lib().d4w_scale_pixels(_lib.ptr(x, "float*"), _lib.ptr(y, "float*"), x.numel(), float(mul), _lib.ptr(ws),
                                               _lib.stream_ptr()), "scale_pixels")
    return y if _is_tensor(img) else _to_host64(y)


def trace2image(trace):
    """|hilbert(trace)| / std_row(trace), min-max scaled to 0..255 (reference: improcess.py:44-63).  The envelope, the
    division by the row sigma and the scaling all run on the GPU (d4w_hilbert mode 3, d4w_scale_pixels)."""
    torch = _torch()
    x = _to_device(trace)
    env = _rows.envelope_over_std(x)
    ws = torch.empty(2, dtype=torch.int32, device=x.device)
    with torch.cuda.device(x.device.index):
        _lib.check(_lib.lib().d4w_scale_pixels(_lib.ptr(env, "float*"), _lib.ptr(env, "float*"), env.numel(), 255.0, _lib.ptr(ws),
                                               _lib.stream_ptr()), "trace2image")
    return env if _is_tensor(trace) else _to_host64(env)


def angle_fromspeed(c0, fs, dx, selected_channels):
    """Angle (degrees) between sound-speed lines and the time axis of the t-x image (reference: improcess.py:66-95)."""
    ratio = c0 / (fs * dx * selected_channels[2])
    print('Detection speed ratio: ', ratio)
    theta_c0 = np.arctan(ratio) * 180 / np.pi
    print('Angle: ', theta_c0)
    return theta_c0


def _gabor_kernel(ksize, sigma, theta, lambd, gamma, psi):
    """cv2.getGaborKernel((ksize, ksize), sigma, theta, lambd, gamma, psi, CV_64F) restated (OpenCV imgproc/src/gabor.cpp):
    half-width xmax = ksize // 2, so the kernel is (2 * (ksize // 2) + 1) square -- 101 x 101 for ksize = 100 -- and
    kernel[ymax - y, xmax - x] = exp(-(xr^2 / 2 sx^2 + yr^2 / 2 sy^2)) * cos(2 pi xr / lambd + psi)."""
    sigma_x, sigma_y = sigma, sigma / gamma
    xmax = ymax = ksize // 2
    c, s = np.cos(theta), np.sin(theta)
    y, x = np.mgrid[-ymax:ymax + 1, -xmax:xmax + 1].astype(np.float64)
    xr = x * c + y * s
    yr = -x * s + y * c
    v = np.exp(-0.5 / sigma_x ** 2 * xr * xr - 0.5 / sigma_y ** 2 * yr * yr) * np.cos(2 * np.pi / lambd * xr + psi)
    return np.ascontiguousarray(v[::-1, ::-1])


def gabor_filt_design(theta_c0, plot=False):
    """Pair of oriented 101 x 101 Gabor kernels for lines along the sound speed (reference: improcess.py:98-140).
    Host side (10 201 values each)."""
    ksize, sigma, lambd, gamma = 100, 4, 20, 0.15
    theta = np.pi / 2 + np.deg2rad(theta_c0)
    gabor_filtup = _gabor_kernel(ksize, sigma, theta, lambd, gamma, 0.0)
    gabor_filtdown = np.flipud(gabor_filtup)
    if plot:
        import matplotlib.pyplot as plt
        plt.figure(figsize=(6, 4))
        for i, g in enumerate((gabor_filtup, gabor_filtdown)):
            plt.subplot(121 + i)
            plt.imshow(g, origin='lower', cmap='RdBu_r', vmin=-1, vmax=1, aspect='equal')
            plt.xlabel('Time indices')
            plt.colorbar(orientation='horizontal')
        plt.tight_layout()
        plt.show()
    return gabor_filtup, gabor_filtdown


def _resize(x, oh, ow):
    torch = _torch()
    ih, iw = x.shape
    out = torch.empty((oh, ow), dtype=torch.float32, device=x.device)
    tmp = torch.empty((ih, ow), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device.index):
        _lib.check(_lib.lib().d4w_resize_aa(_lib.ptr(x, "float*"), ih, iw, _lib.ptr(out, "float*"), oh, ow, _lib.ptr(tmp, "float*"),
                                            _lib.stream_ptr()), "binning")
    return out


def binning(image, ft, fx):
    """Resize an image by ft along time (axis 1) and fx along distance (axis 0) -- torchvision
    `Resize((int(H * fx), int(W * ft)))`, i.e. bilinear interpolation with antialiasing (reference: improcess.py:395-421).
    A boolean image comes back boolean (torchvision resizes it in float32 and casts back: any contributing pixel set)."""
    is_bool = (image.dtype == np.bool_) if isinstance(image, np.ndarray) else (_is_tensor(image) and str(image.dtype) == "torch.bool")
    x = _dev2d(image)
    oh, ow = int(x.shape[0] * fx), int(x.shape[1] * ft)
    y = _resize(x, oh, ow)
    if is_bool:
        y = y != 0
        return y if _is_tensor(image) else y.cpu().numpy()
    return y if _is_tensor(image) else _to_host64(y)


def filter2D(src, ddepth, kernel, in_threshold=None, out_threshold=None, border="reflect101"):
    """GPU stand-in for `cv2.filter2D(src, cv2.CV_64F, kernel)` as scripts/main_gabordetect.py:109,135 call it:
    correlation (no kernel flip), anchor at the kernel centre, BORDER_REFLECT_101.  `ddepth` is accepted and ignored
    (ndarray in -> float64 out, tensor in -> float32 out).  in_threshold / out_threshold binarise the source / the result
    on the fly (`src > thr`), which lets the script's two threshold steps run inside the same kernel."""
    torch = _torch()
    x = _dev2d(src)
    K = torch.from_numpy(np.ascontiguousarray(kernel, dtype=np.float32)).to(x.device)
    if K.ndim != 2 or x.ndim != 2:
        raise ValueError("filter2D expects a 2-D image and a 2-D kernel")
    out = torch.empty_like(x)
    nan = float("nan")
    with torch.cuda.device(x.device.index):
        _lib.check(_lib.lib().d4w_filter2d(_lib.ptr(x, "float*"), x.shape[0], x.shape[1], _lib.ptr(K, "float*"), K.shape[0], K.shape[1],
                                           nan if in_threshold is None else float(in_threshold),
                                           nan if out_threshold is None else float(out_threshold),
                                           0 if border == "reflect101" else 1, _lib.ptr(out, "float*"), _lib.stream_ptr()), "filter2D")
    if _is_tensor(src):
        return out
    return _to_host64(out)


def apply_smooth_mask(array, mask, sigma=1.5):
    """array * mask (reference: improcess.py:424-454 -- the Gaussian-smoothed mask is computed there but the product uses
    the raw mask, :452).  When `mask` is smaller than `array` by an integer factor it is first resized to the array's
    shape like `binning(mask, f, f)` does, fused with the product (d4w_mask_upsample_mul)."""
    torch = _torch()
    x = _to_device(array)
    m = _dev2d(mask)
    nx, ns = x.shape
    out = torch.empty_like(x)
    with torch.cuda.device(x.device.index):
        _lib.check(_lib.lib().d4w_mask_upsample_mul(_lib.ptr(x, "float*"), nx, ns, _lib.ptr(m, "float*"), m.shape[0], m.shape[1],
                                                    _lib.ptr(out, "float*"), _lib.ffi.NULL, _lib.stream_ptr()), "apply_smooth_mask")
    return out if _is_tensor(array) else _to_host64(out)


def gabor_detect(trf_fk, fs, dx, selected_channels, c0=1500., bin_factor=10, threshold=9100., threshold2=150.,
                 return_all=False):
    """The image-domain detector of scripts/main_gabordetect.py:78-169 on the device, end to end:
    trace2image -> binning(1/bin_factor) -> filter2D(up) + filter2D(down) -> > threshold -> filter2D pair on the binary image
    -> > threshold2 -> binning(mask, bin_factor) -> trace * mask.  Both Gabor correlations of a step are one pass with the
    kernel up + down (filter2D is linear in the kernel).  Returns the masked trace (and, with return_all, the intermediate
    images as a dict)."""
    torch = _torch()
    x = _to_device(trf_fk)
    nx, ns = x.shape
    image = trace2image(x)
    imagebin = binning(image, 1 / bin_factor, 1 / bin_factor)
    theta_c0 = np.arctan(c0 / (fs * dx * selected_channels[2])) * 180 / np.pi
    up, down = gabor_filt_design(theta_c0)
    kpair = up + down
    fimage = filter2D(imagebin, None, kpair)
    mask = filter2D(fimage, None, kpair, in_threshold=threshold, out_threshold=threshold2)      # float 0 / 1
    out = torch.empty_like(x)
    with torch.cuda.device(x.device.index):
        _lib.check(_lib.lib().d4w_mask_upsample_mul(_lib.ptr(x, "float*"), nx, ns, _lib.ptr(mask, "float*"), mask.shape[0], mask.shape[1],
                                                    _lib.ptr(out, "float*"), _lib.ffi.NULL, _lib.stream_ptr()), "gabor mask")
    res = out if _is_tensor(trf_fk) else _to_host64(out)
    if not return_all:
        return res
    conv = (lambda t: t) if _is_tensor(trf_fk) else (lambda t: t.to(torch.float64).cpu().numpy())
    return res, {"image": conv(image), "imagebin": conv(imagebin), "fimage": conv(fimage), "mask": conv(mask) != 0}
