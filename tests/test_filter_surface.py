"""Filter-level drop-in: the reference's UNMODIFIED ffmpeg filter source (Transform360/vf_transform360.c), compiled
against the libavfilter stand-in oracle/ffshim and driven by oracle/ff_driver.c, linked with
  * the reference library (oracle/_ref/libt360ref.so + cv2)            -> variant "ref"
  * the product          (transform360_b200/lib/libTransform360.so)    -> variant "b200".
Same filter object code, same option strings; only libTransform360 differs.  Both .so files are prebuilt where
/root/reference exists (make -C oracle ref) and travel to the GPU box."""
import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import ff_harness as ff

needs_filters = pytest.mark.skipif(not (ff.available("ref") and ff.available("b200")), reason="oracle/_ref filter objects not built")


@needs_filters
def test_option_table_and_config_output_through_the_product_headers():
    """No GPU needed: defaults, named constants, ranges and the output-size rules of config_output (vf:167-304)."""
    f = ff.Filter("b200", "cube_edge_length=70", 512, 256)  # rounded down to a multiple of 16 (vf:213) -> 3*64 x 2*64
    assert (f.out_w, f.out_h) == (192, 128)
    f = ff.Filter("b200", "input_stereo_format=TB:cube_edge_length=100", 512, 512)  # TB output doubles the height (vf:293)
    assert (f.out_w, f.out_h) == (288, 384)
    f = ff.Filter("b200", "max_cube_edge_length=1000", 7680, 3840)  # edge = in_w / 4 capped, multiple of 16 (vf:198-213)
    assert (f.out_w, f.out_h) == (2976, 1984)
    f = ff.Filter("b200", "output_layout=eac_32:w=300:h=200", 512, 256)  # non-cubemap layouts take w/h (vf:224-291)
    assert (f.out_w, f.out_h) == (300, 200)
    f = ff.Filter("b200", "output_layout=cubemap_23_offcenter:cube_edge_length=64", 512, 256)
    assert (f.out_w, f.out_h) == (128, 192)
    for bad in ("interpolation_alg=bogus", "num_vertical_segments=1", "nonexistent=1"):
        with pytest.raises(ValueError):
            ff.Filter("b200", bad, 512, 256)
    # the two builds agree on every size decision
    for args, dims in (("cube_edge_length=333", (1920, 960)), ("input_stereo_format=LR:cube_edge_length=96", (2048, 256))):
        a, b = ff.Filter("ref", args, *dims), ff.Filter("b200", args, *dims)
        assert (a.out_w, a.out_h) == (b.out_w, b.out_h)


@needs_filters
@pytest.mark.gpu
@pytest.mark.parametrize("args,dims", [
    ("cube_edge_length=64:interpolation_alg=cubic:enable_low_pass_filter=0", (512, 256)),
    ("cube_edge_length=80:num_vertical_segments=15:num_horizontal_segments=8", (960, 480)),                    # filter defaults: cubic + low-pass
    ("input_stereo_format=TB:output_layout=eac_32:w=192:h=128:interpolation_alg=lanczos4:num_vertical_segments=7", (512, 512)),
    ("cube_edge_length=64:yaw=33:pitch=-12.5:roll=7:interpolation_alg=linear:enable_low_pass_filter=0", (512, 256)),
    ("cube_edge_length=48:width_scale_factor=2:height_scale_factor=2:enable_low_pass_filter=0", (512, 256)),
    ("input_layout=cubemap_32:output_layout=equirect:w=512:h=256:enable_low_pass_filter=0", (384, 256)),
])
def test_unmodified_reference_filter_gives_identical_frames_with_either_library(args, dims):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("needs the B200 box")
    w, h = dims
    planes = [co.noise_plane(w, h, 0, 3), co.noise_plane((w + 1) // 2, (h + 1) // 2, 1, 3), co.noise_plane((w + 1) // 2, (h + 1) // 2, 2, 3)]
    ref, b200 = ff.Filter("ref", args, w, h), ff.Filter("b200", args, w, h)
    assert (ref.out_w, ref.out_h) == (b200.out_w, b200.out_h)
    for frame in range(2):  # the first frame also runs generate_map (vf:346-352)
        want, got = ref.filter(planes), b200.filter(planes)
        for p in range(3):
            assert np.array_equal(got[p], want[p]), f"plane {p} of frame {frame} differs ({args})"
    ref.close()
    b200.close()
