"""Property tests (CPU, hypothesis) of the index arithmetic the crowd input kernels rely on, against the oracle's restatement
of the reference's pad-then-slice code:

  * srgan_crowd_extract_patches reads source pixel (y - P/2 + r, x - P/2 + c) for patch element (r, c) -- mirrored in c when
    flipped -- and writes the padding value where that falls outside the image.  The reference pads the example first and moves
    the centre (crowd/data.py:391-400); the two must agree for every centre, inside or outside, and every image size, including
    images smaller than the patch;
  * srgan_sliding_window_merge lets window (y, x) cover rows [y - P/2, y + P/2) clipped to the image; the reference computes
    start / end offsets per patch (crowd/srgan.py:370-390).  With the positions of ImageSlidingWindowDataset every pixel is
    covered at least once."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import crowd_data_oracle as C


def gather_like_the_kernel(image, label, y, x, patch, flip):
    """The mapping of extract_patches_kernel (sr-gan_b200/csrc/crowd_data.cu), in numpy."""
    H, W = image.shape[:2]
    half = patch // 2
    rows = y - half + np.arange(patch)
    cols = x - half + (patch - 1 - np.arange(patch) if flip else np.arange(patch))
    inside = ((rows >= 0) & (rows < H))[:, None] & ((cols >= 0) & (cols < W))[None, :]
    rr, cc = np.clip(rows, 0, H - 1)[:, None], np.clip(cols, 0, W - 1)[None, :]
    img = np.where(inside[..., None], image[rr, cc], 0).astype(np.uint8)
    lab = np.where(inside, label[rr, cc], 0).astype(np.float32)
    return img, lab


@settings(max_examples=150, deadline=None)
@given(st.integers(3, 40), st.integers(3, 40), st.sampled_from([4, 8, 16, 32]), st.data())
def test_kernel_gather_equals_reference_pad_then_slice(H, W, patch, data):
    rng = np.random.RandomState(H * 1000 + W * 10 + patch)
    image = rng.randint(0, 256, size=(H, W, 3)).astype(np.uint8)
    label = rng.rand(H, W).astype(np.float32)
    y = data.draw(st.integers(0, H - 1))               # allow_padded=True accepts any centre inside the example (:511-513)
    x = data.draw(st.integers(0, W - 1))
    flip = data.draw(st.booleans())
    im_o, lb_o, _ = C.random_horizontal_flip(*C.extract_patch(image, label, label, y, x, patch), flip)
    im_k, lb_k = gather_like_the_kernel(image, label, y, x, patch, flip)
    assert np.array_equal(im_o, im_k) and np.array_equal(lb_o, lb_k)


@settings(max_examples=150, deadline=None)
@given(st.integers(1, 300), st.sampled_from([8, 32, 64, 224]), st.integers(1, 256))
def test_sliding_positions_cover_the_axis(extent, patch, step):
    pos = C.sliding_positions(extent, patch, step)
    half = patch // 2
    assert pos == sorted(set(pos))
    if extent <= half:                                  # the reference yields NO window for such a sliver (:530-537)
        assert pos == []
        return
    assert len(pos) >= 1
    covered = np.zeros(extent, dtype=bool)
    for p in pos:
        lo, hi = max(p - half, 0), min(p + half, extent)
        assert hi > lo                                  # every window overlaps the image
        covered[lo:hi] = True
    if extent >= patch and step <= patch:               # windows no further apart than their width: no pixel left out
        assert covered.all()
    if extent >= patch:
        assert pos[0] == half and pos[-1] == extent - half
    elif extent - half > 0:
        assert pos == [extent - half]                   # one padded window (crowd/data.py:532-533)


@settings(max_examples=60, deadline=None)
@given(st.integers(5, 60), st.integers(5, 60), st.sampled_from([8, 16]), st.integers(3, 20))
def test_merge_clipping_equals_reference_offsets(H, W, patch, step):
    """The gather form of the merge (window covers [y - P/2, y + P/2) clipped) against the oracle's offset arithmetic: same hit
    counts everywhere."""
    half = patch // 2
    ys, xs = C.sliding_positions(H, patch, step), C.sliding_positions(W, patch, step)
    hits_gather = np.zeros((H, W), dtype=np.int32)
    for y in ys:
        for x in xs:
            hits_gather[max(y - half, 0):min(y + half, H), max(x - half, 0):min(x + half, W)] += 1
    image = np.zeros((H, W, 3), dtype=np.uint8)
    count, label = C.predict_full_example(image, lambda im: (np.ones((im.shape[0], patch, patch), np.float32),
                                                             np.full(im.shape[0], float(patch * patch), np.float32), None),
                                          patch, step, 5)
    # a network that returns label 1 and count P^2 per patch: the merged label is 1 where covered, and the count sums to the
    # number of covered pixels
    assert np.array_equal(label > 0, hits_gather > 0)
    assert float(count) == float((hits_gather > 0).sum())
