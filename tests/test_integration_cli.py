"""The drop-in itself (VERDICT r01 N1 / SURVEY T3): the reference's own command-line encoder built twice from
/root/reference by integration/Makefile -- unmodified (CPU lookahead) and with integration/x265_enable_cuda.patch +
-DENABLE_CUDA=1 (the Lookahead class forwarding to the B200 engine) -- must produce the same frame types, scene cuts,
I/P cost ratios, QPs and the same BITSTREAM, bit for bit."""
import hashlib
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "integration", "_build")


def _bin(kind, depth):
    return os.path.join(BUILD, "x265-%s%d" % (kind, depth))


def test_integration_patch_and_binaries():
    """the patch only hooks the public entry points of Lookahead and leaves the class declaration alone; where the
    reference sources exist the two encoders were built from them (by __graft_entry__.build) and the CUDA one is linked
    against the product libraries"""
    patch = open(os.path.join(ROOT, "integration", "x265_enable_cuda.patch")).read()
    assert "a/source/encoder/slicetype.h" not in patch and "a/source/encoder/encoder.cpp" not in patch
    for hook in ("cudaLookaheadCreate", "cudaLookaheadAddPicture", "cudaLookaheadGetDecided", "cudaLookaheadEstimatedPictureCost",
                 "cudaLookaheadFlush", "cudaLookaheadDestroy", "cudaLookaheadFindSliceType", "option(ENABLE_CUDA"):
        assert hook in patch, hook
    if not os.path.isdir("/root/reference/source"):
        pytest.skip("reference sources not on this box")
    for depth in (8, 10, 12):
        assert os.path.exists(_bin("cpu", depth)) and os.path.exists(_bin("cuda", depth))
        needed = subprocess.run(["readelf", "-d", _bin("cuda", depth)], stdout=subprocess.PIPE, text=True).stdout
        assert "libx265la.so" in needed      # which in turn needs libx265cu.so, the CUDA engine
        needed = subprocess.run(["readelf", "-d", _bin("cpu", depth)], stdout=subprocess.PIPE, text=True).stdout
        assert "libx265la.so" not in needed


def _encode(binary, y4m, out, extra, env=None):
    csv = out + ".csv"
    cmd = [binary, "--input", y4m, "--frame-threads", "1", "--csv", csv, "--csv-log-level", "2", "--log-level", "error",
           "-o", out] + extra
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=e, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:]
    rows = []
    with open(csv) as f:
        head = [h.strip() for h in f.readline().split(",")]
        keep = [head.index(k) for k in ("Encode Order", "Type", "POC", "QP", "Bits", "Scenecut", "I/P cost ratio")]
        for line in f:
            cells = [c.strip() for c in line.split(",")]
            if len(cells) > max(keep) and cells[0].isdigit():
                rows.append([cells[i] for i in keep])
    return hashlib.md5(open(out, "rb").read()).hexdigest(), rows


CLI_CASES = [
    ("medium_360p", 8, 640, 360, 60, dict(cuts=(33,)), ["--preset", "medium", "--pools", "4", "--lookahead-slices", "0"]),
    ("medium_720p_slices", 8, 1280, 720, 24, dict(cuts=(11,)), ["--preset", "medium", "--pools", "4"]),
    ("veryfast_nopool", 8, 640, 360, 50, dict(cuts=(20,)), ["--preset", "veryfast", "--pools", "none", "--no-wpp"]),
    ("vbv_10bit", 10, 640, 360, 50, dict(cuts=(27,)), ["--preset", "fast", "--pools", "4", "--lookahead-slices", "0", "--bitrate", "1500",
                                                          "--vbv-bufsize", "2000", "--vbv-maxrate", "2000"]),
    ("b8_la40_fades_10bit", 10, 640, 360, 70, dict(cuts=(), fades=[(20, 12, 0.3), (45, 10, 1.0)]),
     ["--preset", "medium", "--pools", "16", "--lookahead-slices", "0", "--bframes", "8", "--rc-lookahead", "40", "--weightb"]),
    ("qg8_cqp", 8, 640, 360, 40, dict(cuts=(19,)), ["--preset", "faster", "--pools", "4", "--qg-size", "8", "--aq-mode", "3", "--crf", "24"]),
    # --fades (10 fps, so that the 16-frame fade-ins last "at least one second"): the frame that ends a fade-in becomes a keyframe
    ("fades_10fps", 8, 640, 360, 70, dict(cuts=(), envelope=[(0, 1.0), (6, 1.0), (12, 0.12), (16, 0.12), (32, 1.0), (44, 1.0), (47, 0.3), (62, 1.0)]),
     ["--preset", "medium", "--pools", "4", "--lookahead-slices", "0", "--fades"]),
    # --hist-scenecut (8-bit) and two temporal layers
    ("hist_scenecut", 8, 640, 360, 60, dict(cuts=(17, 41), envelope=[(0, 1.0), (16, 1.0), (17, 0.5), (40, 0.5), (41, 0.95), (59, 0.95)]),
     ["--preset", "medium", "--pools", "4", "--lookahead-slices", "0", "--hist-scenecut"]),
    # --hme: the lookahead's level-0 / level-1 searches on the GPU (hex, umh), the main encoder's level 2 fed by the mirrored lowres MVs
    ("hme_544p", 8, 960, 544, 24, dict(cuts=(11,)), ["--preset", "medium", "--pools", "4", "--lookahead-slices", "0", "--hme"]),
    # aq-mode 4 (edge): the qp offsets the frame encoder quantises with come from the GPU's edge map
    ("aq4_edge", 8, 640, 360, 40, dict(cuts=(19,)), ["--preset", "medium", "--pools", "4", "--lookahead-slices", "0", "--aq-mode", "4"]),
    ("temporal_layers_2", 8, 640, 360, 50, dict(cuts=(23,)),
     ["--preset", "medium", "--pools", "4", "--lookahead-slices", "0", "--bframes", "7", "--temporal-layers", "2"]),
    # 3 and 5 temporal layers: random-access mini-GOPs of 4 / 16 pictures, split at the scene cuts and at the end of the stream; the DPB
    # builds its reference picture sets from the positions / layers the GPU lookahead hands back
    ("temporal_layers_3", 8, 640, 360, 46, dict(cuts=(14, 30)), ["--preset", "medium", "--pools", "4", "--lookahead-slices", "0", "--temporal-layers", "3"]),
    ("temporal_layers_5", 10, 640, 360, 70, dict(cuts=(13, 41, 52)),
     ["--preset", "medium", "--pools", "16", "--lookahead-slices", "0", "--rc-lookahead", "35", "--temporal-layers", "5"]),
    # main12: the 12-bit build of the encoder (16-bit samples, 32-bit SATD butterflies on the GPU)
    ("main12_weightb", 12, 640, 360, 40, dict(cuts=(), fades=[(12, 10, 0.3)]), ["--preset", "medium", "--pools", "4", "--lookahead-slices", "0", "--weightb"]),
    # --lookahead-threads: the lookahead gets a pool of its own, whose size selects the reference's batch modes
    ("lookahead_threads", 8, 640, 360, 50, dict(cuts=(23,)), ["--preset", "medium", "--pools", "8", "--lookahead-threads", "2", "--lookahead-slices", "0"]),
]


@pytest.mark.gpu
@pytest.mark.parametrize("name", [c[0] for c in CLI_CASES])
def test_x265_cli_with_cuda_lookahead_is_bit_identical(name, synth, tmp_path):
    name, depth, w, h, n, skw, extra = [c for c in CLI_CASES if c[0] == name][0]
    if not (os.path.exists(_bin("cpu", depth)) and os.path.exists(_bin("cuda", depth))):
        pytest.skip("integration/_build not in this snapshot (needs /root/reference at build time)")
    seq = synth.SynthSequence(w, h, depth=depth, seed=5, n_rects=5, **skw)
    y4m = str(tmp_path / "in.y4m")
    synth.write_y4m(y4m, seq, n, fps=(10, 1) if "--fades" in extra else (30, 1))
    md5_cpu, rows_cpu = _encode(_bin("cpu", depth), y4m, str(tmp_path / "cpu.hevc"), extra)
    md5_gpu, rows_gpu = _encode(_bin("cuda", depth), y4m, str(tmp_path / "gpu.hevc"), extra, env={"X265_CUDA_ASYNC_DEPTH": "8"})
    assert len(rows_cpu) == n and len(rows_gpu) == n
    # Lowres::ipCostRatio is only ever written by Lookahead::scenecut (slicetype.cpp:2998-3003) and never reset when a Frame is
    # recycled (lowres.cpp:337-365): with --hist-scenecut the first-frame check goes through histBasedScenecut instead, and the
    # CSV column of most frames shows whatever an EARLIER frame left in the recycled object.  Not reproducible, not compared.
    skip_ratio = "--hist-scenecut" in extra
    for a, b in zip(rows_cpu, rows_gpu):
        if skip_ratio:
            a, b = a[:6] + a[7:], b[:6] + b[7:]
        assert a == b, "csv row differs:\ncpu %s\ngpu %s" % (a, b)
    assert md5_cpu == md5_gpu, "bitstreams differ although every csv row matches"
