"""The one helper of the reference's util.py that sits on the hot-path boundary."""


def center_crop(img, dst_shape):
    """Symmetric spatial crop by slicing (a view, no copy) -- util.py:92-114.
    Called on the network outputs at train.py:414-417 (192 -> 180)."""
    src_nr, src_nc = img.shape[-2], img.shape[-1]
    dst_nr, dst_nc = dst_shape[-2], dst_shape[-1]
    if (dst_nr != src_nr) or (dst_nc != src_nc):
        r0 = int((src_nr - dst_nr) / 2)
        c0 = int((src_nc - dst_nc) / 2)
        if img.dim() == 4:
            return img[:, :, r0:r0 + dst_nr, c0:c0 + dst_nc]
        elif img.dim() == 3:
            return img[:, r0:r0 + dst_nr, c0:c0 + dst_nc]
        else:
            assert img.dim() == 2
            return img[r0:r0 + dst_nr, c0:c0 + dst_nc]
    return img
