"""CPU checks of the C-ABI boundary: the library builds, loads, exports every symbol
include/yolonano_b200.h declares, agrees with the Python topology, and refuses to run
without a GPU (no CPU fallback).  No compute calls here."""
import ctypes as C
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "yolonano_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ynb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_bound_and_exported(lib):
    from yolo_nano_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 25
    assert sorted(_lib.SIGNATURES.keys()) == declared, "ctypes table and header disagree"
    for name in declared:
        assert getattr(lib, name) is not None


def test_no_oracle_or_torch_ops_in_product_path():
    """The product package must not import the oracle (parity would be void)."""
    for f in (ROOT / "yolo_nano_b200").glob("*.py"):
        src = f.read_text()
        assert "import oracle" not in src and "from oracle" not in src, f.name
        assert "F.conv2d" not in src and "torch.nn.functional" not in src, f.name


def test_conv_table_matches_engine(lib):
    from yolo_nano_b200.topology import conv_table
    for classes in (20, 80):
        table = conv_table(classes)
        assert lib.ynb_num_convs() == len(table) == 77
        for i, spec in enumerate(table):
            assert lib.ynb_conv_name(i).decode() == spec.name
            co, ci, k = C.c_int32(), C.c_int32(), C.c_int32()
            assert lib.ynb_conv_shape(i, classes, 3, C.byref(co), C.byref(ci), C.byref(k)) == 0
            assert (co.value, ci.value, k.value, k.value) == spec.weight_shape(), spec.name


def test_state_dict_covers_table():
    import contextlib, io
    import yolo_nano_b200 as pkg
    with contextlib.redirect_stdout(io.StringIO()):
        m = pkg.YOLONano(torch.device("cpu"), 320, 20, anchor_size=pkg.MULTI_ANCHOR_SIZE)
    sd = m.state_dict()
    assert len(sd) == 469 and len(list(m.parameters())) == 247
    for spec in pkg.conv_table(20):
        assert tuple(sd[spec.name + ".weight"].shape) == spec.weight_shape()
        assert ((spec.name + ".bias") in sd) == spec.conv_bias
        if spec.bn:
            assert spec.bn + ".running_var" in sd
    fw = m.fused_weights()
    assert set(fw) == {s.name for s in pkg.conv_table(20)}


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a GPU-less machine")
def test_fails_loudly_without_gpu(lib):
    import contextlib, io
    import yolo_nano_b200 as pkg
    from yolo_nano_b200 import _lib
    cfg = _lib.YnbConfig()
    cfg.abi_version = _lib.YNB_ABI_VERSION
    cfg.input_size, cfg.num_classes, cfg.num_anchors, cfg.max_batch = 320, 20, 3, 1
    h = C.c_void_p()
    rc = lib.ynb_create(C.byref(cfg), C.byref(h))
    assert rc == 2 and not h.value          # YNB_ERR_NO_DEVICE
    assert b"no CUDA device" in lib.ynb_last_error(None)
    with contextlib.redirect_stdout(io.StringIO()):
        m = pkg.YOLONano(torch.device("cpu"), 320, 20, anchor_size=pkg.MULTI_ANCHOR_SIZE).eval()
    with pytest.raises(pkg.EngineError):
        m(torch.zeros(1, 3, 320, 320))


def test_reference_api_surface():
    import contextlib, copy, io
    import yolo_nano_b200 as pkg
    with contextlib.redirect_stdout(io.StringIO()):
        m = pkg.YOLONano(torch.device("cpu"), 416, 80, anchor_size=pkg.MULTI_ANCHOR_SIZE_COCO,
                         conf_thresh=0.1, nms_thresh=0.45, diou_nms=True)
    assert m.stride == [8, 16, 32] and m.num_anchors == 3 and m.input_size == 416
    m.set_grid(320)
    assert m.input_size == 320
    g, s, a = m.create_grid(320)
    assert g.shape == (1, 2100, 1, 2) and s.shape == (1, 2100, 3, 2) and a.shape == (1, 2100, 3, 2)
    m2 = copy.deepcopy(m)            # ModelEMA does this (utils/misc.py:70)
    assert len(m2.state_dict()) == 469
    m.trainable = True
    with pytest.raises(ValueError):                    # the training branch needs target (models/yolo_nano.py:333)
        m(torch.zeros(1, 3, 320, 320))
    with pytest.raises(pkg.EngineError):               # and a CUDA device: no CPU fallback
        m(torch.zeros(1, 3, 320, 320), target=torch.zeros(1, 2100, 11))
    with pytest.raises(SystemExit):
        with contextlib.redirect_stdout(io.StringIO()):
            pkg.YOLONano(torch.device("cpu"), 416, 80, anchor_size=pkg.MULTI_ANCHOR_SIZE_COCO, backbone="0.5x")


def test_create_grid_values_match_reference_layout():
    """create_grid (models/yolo_nano.py:86-112): values, not only shapes — grid_xy = (col, row) row-major per
    level, strides and anchors repeated per cell — against the oracle's restatement."""
    import contextlib
    import io
    import yolo_nano_b200 as pkg
    from oracle import yolo_nano_oracle as O
    with contextlib.redirect_stdout(io.StringIO()):
        m = pkg.YOLONano(torch.device("cpu"), 320, 20, anchor_size=pkg.MULTI_ANCHOR_SIZE)
    want = O.grid_tensors(320, pkg.MULTI_ANCHOR_SIZE)
    for got, w_ in zip(m.create_grid(320), want):
        assert got.shape == w_.shape and torch.equal(got.cpu(), w_)
    m.set_grid(416)
    for got, w_ in zip((m.grid_cell, m.stride_tensor, m.all_anchors_wh), O.grid_tensors(416, pkg.MULTI_ANCHOR_SIZE)):
        assert torch.equal(got.cpu(), w_)
