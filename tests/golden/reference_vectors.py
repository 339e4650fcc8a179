"""Golden vectors transcribed from the reference's own tests and doc examples.

Each case is a sequence of *segments* (model, symbols, per-symbol parameters)
coded with one coder, plus the exact `uint32` words the reference asserts (or
prints in its doc comments).  `src` cites /root/reference file:line.

ANS segments are listed in *decode* order; the encoder pushes them with
`encode_reverse` in reverse segment order, as the reference's tests do.

Model specs:
  ("qgauss", min, max, mean|None, std|None)       QuantizedGaussian
  ("cat", probabilities|None, kwargs)             Categorical(**kwargs)
Per-symbol parameters (means/stds or an (m,n) probability matrix) go in `params`.
"""
import numpy as np

f32, f64 = np.float32, np.float64


def A(x, dt):
    return np.array(x, dtype=dt)


def _cases():
    out = []

    def add(id, src, coder, segments, words):
        out.append(dict(id=id, src=src, coder=coder, segments=segments,
                        words=np.array(words, dtype=np.uint32)))

    # --- G1/G2: src/lib.rs:112-131,209-244; tests/python/test_constriction.py:6-55
    syms = A([23, -15, 78, 43, -69], np.int32)
    means = [35.2, -1.7, 30.1, 71.2, -75.1]
    stds = [10.1, 25.3, 23.8, 35.4, 3.9]
    add("G1_ans_gauss_params_f64", "src/lib.rs:131; tests/python/test_constriction.py:32-45", "ans",
        [(("qgauss", -100, 100, None, None), syms, (A(means, f64), A(stds, f64)))], [1109163715, 757457])
    add("G2_range_gauss_params_f64", "src/lib.rs:244; tests/python/test_constriction.py:6-19", "range",
        [(("qgauss", -100, 100, None, None), syms, (A(means, f64), A(stds, f64)))], [473034731, 2276733146])

    # --- G3/G4: tests/python/test_docexamples_f32.py:7-42 (the BASELINE model)
    msg = A([6, 10, -4, 2, 5, 2, 1, 0, 2], np.int32)
    add("G3_ans_gauss_iid", "tests/python/test_docexamples_f32.py:7-21", "ans",
        [(("qgauss", -50, 50, 3.2, 9.6), msg, ())], [3114258274, 357938615])
    add("G4_range_gauss_iid", "tests/python/test_docexamples_f32.py:28-42", "range",
        [(("qgauss", -50, 50, 3.2, 9.6), msg, ())], [2682585243, 513522013])

    # --- G5/G6: test_docexamples.py:730-758 (+ _f32.py:733-760)
    syms = A([12, 15, 4, -2, 18, 5], np.int32)
    add("G5_ans_gauss_iid_100", "tests/python/test_docexamples.py:730-738", "ans",
        [(("qgauss", -100, 100, 12.6, 7.3), syms, ())], [745994372, 25704])
    means = [13.2, 17.9, 7.3, -4.2, 25.1, 3.2]
    stds = [3.2, 4.7, 5.2, 3.1, 6.3, 2.9]
    add("G6_ans_gauss_params_f64", "tests/python/test_docexamples.py:744-758", "ans",
        [(("qgauss", -100, 100, None, None), syms, (A(means, f64), A(stds, f64)))], [2051958011, 1549])
    add("G6_ans_gauss_params_f32", "tests/python/test_docexamples_f32.py:747-760", "ans",
        [(("qgauss", -100, 100, None, None), syms, (A(means, f32), A(stds, f32)))], [2051912079, 1549])

    # --- G7/G8: test_docexamples.py:369-384, 558-573 (+ f32 variants give the same words)
    syms = A([12, -13, 25], np.int32)
    means, stds = [10.3, -4.7, 20.5], [5.2, 24.2, 3.1]
    for dt, tag, file in ((f64, "f64", "test_docexamples.py"), (f32, "f32", "test_docexamples_f32.py")):
        add(f"G7_ans_gauss3_{tag}", f"tests/python/{file}:369-384", "ans",
            [(("qgauss", -100, 100, None, None), syms, (A(means, dt), A(stds, dt)))], [597775281, 3])
        add(f"G8_range_gauss3_{tag}", f"tests/python/{file}:558-573", "range",
            [(("qgauss", -100, 100, None, None), syms, (A(means, dt), A(stds, dt)))], [2655472005])

    # --- G9: test_docexamples.py:764-775, _f32.py:767-777, test_lazy_f{32,64}.py
    syms = A([0, 3, 2, 3, 2, 0, 2, 1], np.int32)
    p = [0.2, 0.4, 0.1, 0.3]
    for kw, kt in (({"perfect": False}, "fast"), ({"lazy": True}, "lazy")):
        add(f"G9_ans_cat_f64_{kt}", "tests/python/test_docexamples.py:764-775; test_lazy_f64.py:388-399", "ans",
            [(("cat", A(p, f64), kw), syms, ())], [488222996, 175])
        add(f"G9_ans_cat_f32_{kt}", "tests/python/test_docexamples_f32.py:767-777; test_lazy_f32.py:416-427", "ans",
            [(("cat", A(p, f32), kw), syms, ())], [2484720979, 175])

    # --- G10/G11: test_docexamples.py:356-366, 545-555 (+ f32, lazy)
    syms = A([0, 2, 1, 2, 0, 2, 0, 2, 1], np.int32)
    p = [0.1, 0.6, 0.3]
    for kw, kt in (({"perfect": False}, "fast"), ({"lazy": True}, "lazy")):
        add(f"G10_ans_cat3_f64_{kt}", "tests/python/test_docexamples.py:356-366; test_lazy_f64.py:181-191", "ans",
            [(("cat", A(p, f64), kw), syms, ())], [1276728145, 172])
        add(f"G10_ans_cat3_f32_{kt}", "tests/python/test_docexamples_f32.py:358-368; test_lazy_f32.py:207-217", "ans",
            [(("cat", A(p, f32), kw), syms, ())], [1276732052, 172])
        add(f"G11_range_cat3_f64_{kt}", "tests/python/test_docexamples.py:545-555; test_lazy_f64.py:292-302", "range",
            [(("cat", A(p, f64), kw), syms, ())], [369323576])
        add(f"G11_range_cat3_f32_{kt}", "tests/python/test_docexamples_f32.py:547-557; test_lazy_f32.py:319-329", "range",
            [(("cat", A(p, f32), kw), syms, ())], [369323598])

    # --- G12/G13: per-symbol categorical families, test_docexamples.py:387-400, 576-589 (+ f32, lazy)
    syms = A([3, 1], np.int32)
    pm = [[0.1, 0.2, 0.3, 0.1, 0.3], [0.3, 0.2, 0.2, 0.2, 0.1]]
    for kw, kt in (({"perfect": False}, "fast"), ({"lazy": True}, "lazy")):
        add(f"G12_ans_catfam_f64_{kt}", "tests/python/test_docexamples.py:387-400; test_lazy_f64.py:195-208", "ans",
            [(("cat", None, kw), syms, (A(pm, f64),))], [45298481])
        add(f"G12_ans_catfam_f32_{kt}", "tests/python/test_docexamples_f32.py:389-402; test_lazy_f32.py:221-235", "ans",
            [(("cat", None, kw), syms, (A(pm, f32),))], [45298482])
        add(f"G13_range_catfam_f64_{kt}", "tests/python/test_docexamples.py:576-589; test_lazy_f64.py:306-319", "range",
            [(("cat", None, kw), syms, (A(pm, f64),))], [2705829254])
        add(f"G13_range_catfam_f32_{kt}", "tests/python/test_docexamples_f32.py:578-591; test_lazy_f32.py:333-346", "range",
            [(("cat", None, kw), syms, (A(pm, f32),))], [2705829510])

    # --- G14: test_docexamples.py:781-794 (+ f32: src/pybindings/stream/model.rs:450)
    syms = A([0, 4, 1], np.int32)
    pm = [[0.3, 0.1, 0.1, 0.3, 0.2], [0.1, 0.4, 0.2, 0.1, 0.2], [0.4, 0.2, 0.1, 0.2, 0.1]]
    for kw, kt in (({"perfect": False}, "fast"), ({"lazy": True}, "lazy")):
        add(f"G14_ans_catfam3_f64_{kt}", "tests/python/test_docexamples.py:781-794; test_lazy_f64.py:405-418", "ans",
            [(("cat", None, kw), syms, (A(pm, f64),))], [104018741])
        add(f"G14_ans_catfam3_f32_{kt}", "tests/python/test_docexamples_f32.py:783-796; test_lazy_f32.py:433-446", "ans",
            [(("cat", None, kw), syms, (A(pm, f32),))], [104018743])

    # --- G16: two models in one stream, test_docexamples.py:90-111 (+ _f32.py:92-113, test_lazy_*.py)
    msg = A([6, 10, -4, 2, 5, 2, 1, 0, 2], np.int32)
    means, stds = [2.3, 6.1, -8.5, 4.1, 1.3], [6.2, 5.3, 3.8, 3.2, 4.7]
    add("G16_range_mixed_f64", "tests/python/test_docexamples.py:90-111", "range",
        [(("qgauss", -50, 50, None, None), msg[:5], (A(means, f64), A(stds, f64))),
         (("cat", A([0.2, 0.5, 0.3], f64), {"perfect": False}), msg[5:], ())], [3176507208])
    add("G16_range_mixed_f32", "tests/python/test_docexamples_f32.py:92-113", "range",
        [(("qgauss", -50, 50, None, None), msg[:5], (A(means, f32), A(stds, f32))),
         (("cat", A([0.2, 0.5, 0.3], f32), {"perfect": False}), msg[5:], ())], [3176507206])
    add("G16_range_mixed_lazy_f64params_f32probs", "tests/python/test_lazy_f64.py:6-27", "range",
        [(("qgauss", -50, 50, None, None), msg[:5], (A(means, f64), A(stds, f64))),
         (("cat", A([0.2, 0.5, 0.3], f32), {"lazy": True}), msg[5:], ())], [3176507208])
    add("G16_range_mixed_lazy_f32", "tests/python/test_lazy_f32.py:33-54", "range",
        [(("qgauss", -50, 50, None, None), msg[:5], (A(means, f32), A(stds, f32))),
         (("cat", A([0.2, 0.5, 0.3], f32), {"lazy": True}), msg[5:], ())], [3176507206])
    return out


ENCODE_CASES = _cases()

# Decode-only goldens: (id, src, coder, model spec, params-or-count, compressed, expected symbols)
DECODE_CASES = [
    ("G15_ans_decode_single_f64", "tests/python/test_docexamples.py:287-297", "ans",
     ("cat", A([0.1, 0.6, 0.3], f64), {"perfect": False}), None, [2514924296, 114], 2),
    ("G15_ans_decode9_f64", "tests/python/test_docexamples.py:300-311", "ans",
     ("cat", A([0.1, 0.6, 0.3], f64), {"perfect": False}), 9, [1441153686, 108], [2, 0, 0, 1, 2, 2, 1, 2, 2]),
    ("G15_ans_decode9_f32", "tests/python/test_docexamples_f32.py:300-311", "ans",
     ("cat", A([0.1, 0.6, 0.3], f32), {"perfect": False}), 9, [2514924296, 114], [2, 0, 0, 1, 2, 2, 1, 2, 2]),
    ("G15_ans_decode9_f64_lazy", "tests/python/test_lazy_f64.py:140-151", "ans",
     ("cat", A([0.1, 0.6, 0.3], f64), {"lazy": True}), 9, [1441153686, 108], [2, 0, 0, 1, 2, 2, 1, 2, 2]),
    ("ans_decode_catfam_f64", "tests/python/test_docexamples.py:330-342", "ans",
     ("cat", None, {"perfect": False}), (A([[0.1, 0.2, 0.3, 0.1, 0.3], [0.3, 0.2, 0.2, 0.2, 0.1]], f64),),
     [2142112014, 31], [3, 1]),
    ("ans_decode_catfam_f32", "tests/python/test_docexamples_f32.py:332-344", "ans",
     ("cat", None, {"perfect": False}), (A([[0.1, 0.2, 0.3, 0.1, 0.3], [0.3, 0.2, 0.2, 0.2, 0.1]], f32),),
     [2142112014, 31], [3, 1]),
    ("range_decode_single", "tests/python/test_docexamples.py:592-602", "range",
     ("cat", A([0.1, 0.6, 0.3], f64), {"perfect": False}), None, [3089773345, 1894195597], 2),
    ("G13_range_decode_noninjective", "tests/python/test_docexamples.py:659-671", "range",
     ("cat", None, {"perfect": False}), (A([[0.1, 0.2, 0.3, 0.1, 0.3], [0.3, 0.2, 0.2, 0.2, 0.1]], f64),),
     [2705829535], [3, 1]),
]

# G17: tests/python/test_constriction.py:102-117 -- AnsCoder(data, seal=True) + lazy categoricals
SEAL_DATA = np.array([0x80d14131, 0xdda97c6c, 0x5017a640, 0x01170a3e], dtype=np.uint32)
SEAL_PROBS = np.array([[0.1, 0.7, 0.1, 0.1], [0.2, 0.2, 0.1, 0.5], [0.2, 0.1, 0.4, 0.3]])
SEAL_EXPECT = [([0.1, 0.7, 0.1, 0.1], [0, 0, 2]), ([0.09, 0.71, 0.1, 0.1], [1, 0, 0])]

# G18: src/stream/stack.rs:1250-1291, src/stream/queue.rs:1133-1173 -- word counts,
# QuantizedGaussian(-127,127,3.2,5.1) i.i.d.
LENGTH_CASES = [([5], 1), ([2, 8], 1), (list(range(10)), 2), (list(range(-10, 10)), 4)]
