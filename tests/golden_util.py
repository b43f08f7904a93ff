"""Loads tests/golden/reference_vectors.json and checks an implementation against it.

``api`` is any object exposing the numpy-level functions of ``oracle`` (radix_sort,
insertion_sort, sort, stable_sort, sort_by_key, stable_sort_by_key, scan, reduce,
accumulate): the oracle module itself, or tests/gpu_api.py which routes the same calls
through the C-ABI of the CUDA library.
"""
from __future__ import annotations

import json
import os

import numpy as np

NP = {
    "char": np.int8, "uchar": np.uint8, "short": np.int16, "ushort": np.uint16,
    "int": np.int32, "uint": np.uint32, "long": np.int64, "ulong": np.uint64,
    "float": np.float32, "double": np.float64,
}

_HERE = os.path.dirname(os.path.abspath(__file__))


def load_cases():
    with open(os.path.join(_HERE, "golden", "reference_vectors.json")) as f:
        return json.load(f)["cases"]


def case_id(c):
    extra = "desc" if c.get("descending") else ""
    return f"{c['ref']}-{c['fn']}-{c.get('dtype', '')}-{c.get('op', '')}{extra}".replace(" ", "")


def _expand_input(c):
    dt = NP[c["dtype"]]
    if "gen" in c:
        g = c["gen"]
        if g["kind"] == "iota":
            return np.arange(g["start"], g["start"] + g["n"]).astype(dt)
        if g["kind"] == "fill":
            return np.full(g["n"], g["value"]).astype(dt)
        return _expand_gen(g, dt)
    return np.array(c.get("input", []), dtype=dt)


def _std_accumulate(x, init_np):
    """std::accumulate(data.begin(), data.end(), 0): `acc = acc + x` evaluated in the promoted
    type, then converted back to int (truncation) every step (test_accumulate.cpp:258-268)."""
    acc = init_np.type(0)
    for v in x:
        acc = init_np.type(x.dtype.type(acc) + v)
    return acc


def check_case(c, api):
    fn = c["fn"]
    desc = bool(c.get("descending", False))
    if fn in ("radix_sort", "insertion_sort", "sort", "stable_sort"):
        x = _expand_input(c)
        if "sub_range" in c:
            lo, hi = c["sub_range"]
            got = api.sort_sub_range(fn, x, lo, hi, desc) if hasattr(api, "sort_sub_range") else _slice_sort(api, fn, x, lo, hi, desc)
        elif c.get("host_range") and hasattr(api, "sort_host"):
            got = api.sort_host(x, desc)
        else:
            got = getattr(api, fn)(x, desc)
        exp = np.array(c["expected"], dtype=x.dtype)
        assert got.dtype == x.dtype
        np.testing.assert_array_equal(got, exp, err_msg=c["ref"])  # value equality (== like CHECK_RANGE_EQUAL)
        return
    if fn in ("radix_sort_by_key", "sort_by_key", "stable_sort_by_key"):
        kd, vd, vw = NP[c["dtype"]], NP[c["value_dtype"]], c["value_width"]
        if "gen" in c:
            g = c["gen"]
            n = g["n"]
            if g["kind"] == "reversed_keys_with_payload":
                keys = (n - np.arange(n)).astype(kd)
                vals = np.repeat(((n - np.arange(n)) / 2.0).astype(vd)[:, None], vw, axis=1)
            elif g["kind"] == "mid_stability":
                keys = (-np.arange(n)).astype(kd)
                vals = (-np.arange(n)).astype(vd)[:, None].copy()
                keys[n // 2] = keys[n - 2] = keys[n - 1] = -256
                vals[n // 2], vals[n - 2], vals[n - 1] = 3, 1, 2
            else:
                raise ValueError(g["kind"])
        else:
            keys = np.array(c["keys"], dtype=kd)
            vals = np.array(c["values"], dtype=vd).reshape(-1, vw)
        if fn == "radix_sort_by_key":
            gk, gv = api.radix_sort(keys, desc, vals)
        else:
            gk, gv = getattr(api, fn)(keys, vals, desc)
        gv = np.asarray(gv).reshape(-1, vw)
        if "expected_keys" in c:
            np.testing.assert_array_equal(gk, np.array(c["expected_keys"], dtype=kd), err_msg=c["ref"])
            np.testing.assert_array_equal(gv, np.array(c["expected_values"], dtype=vd).reshape(-1, vw), err_msg=c["ref"])
        elif "expected_head_keys" in c:
            h = len(c["expected_head_keys"])
            np.testing.assert_array_equal(gk[:h], np.array(c["expected_head_keys"], dtype=kd))
            np.testing.assert_array_equal(gv[:h, 0], np.array(c["expected_head_values"], dtype=vd))
            assert np.all(gk[:-1] <= gk[1:])
        else:  # reversed keys: sorted keys ascending, payload follows its key
            order = np.argsort(keys, kind="stable")
            np.testing.assert_array_equal(gk, keys[order])
            np.testing.assert_array_equal(gv, vals[order])
        return
    if fn == "scan":
        x = _expand_input(c)
        if "expected_gen" in c:
            t = x.copy()
            if c["expected_gen"].endswith("init10"):
                t[0] = 10
            with np.errstate(over="ignore"):
                exp = np.multiply.accumulate(t, dtype=x.dtype)  # std::partial_sum(multiplies), int wrap-around
        else:
            exp = np.array(c["expected"], dtype=x.dtype)
        variants = [False, True] if c.get("also_in_place") else [False]
        for in_place in variants:
            got = api.scan(x, c["op"], bool(c["exclusive"]), c["init"], in_place=in_place) if _accepts(api.scan, "in_place") \
                else api.scan(x, c["op"], bool(c["exclusive"]), c["init"])
            if "rel_tol" in c:
                np.testing.assert_allclose(got, exp, rtol=c["rel_tol"], err_msg=c["ref"])
            else:
                np.testing.assert_array_equal(got, exp, err_msg=c["ref"])
        return
    if fn == "reduce":
        x = _expand_input(c)
        if "sub_range" in c:
            lo, hi = c["sub_range"]
            x = x[lo:hi]
        rdt = NP[c.get("result_dtype", c["dtype"])]
        if c.get("untouched"):
            got = api.reduce_into(x, c["op"], rdt, rdt(c["expected"])) if hasattr(api, "reduce_into") else rdt(c["expected"])
        elif c.get("result_on_device") and hasattr(api, "reduce_to_device"):
            got = api.reduce_to_device(x, c["op"], rdt)
        else:
            got = api.reduce(x, c["op"], rdt)
        assert rdt(got) == rdt(c["expected"]), (c["ref"], got, c["expected"])
        return
    if fn == "accumulate":
        x = _expand_input(c)
        adt = NP[c["init_dtype"]]
        got = api.accumulate(x, adt(c["init"]), c["op"], op_dtype=x.dtype, acc_dtype=adt)
        if "expected_gen" in c:
            exp = _std_accumulate(x, np.dtype(adt))
        else:
            exp = adt(c["expected"])
        assert adt(got) == exp, (c["ref"], got, exp)
        return
    if fn in ("copy_if", "transform_if"):
        x = _expand_input(c)
        pred = tuple(c["pred"])
        if _accepts(api.transform_if, "fill"):  # GPU adapter: output buffer pre-filled, tail must stay untouched
            out, count = api.transform_if(x, c.get("function", "identity"), pred, fill=c["fill"])
        else:
            sel = api.transform_if(x, c.get("function", "identity"), pred)
            count = len(sel)
            out = np.full(x.size, c["fill"], dtype=x.dtype)
            out[:count] = sel
        assert count == c["count"], (c["ref"], count)
        np.testing.assert_array_equal(out, np.array(c["expected"], dtype=x.dtype), err_msg=c["ref"])  # untouched tail included
        return
    if fn in ("count", "count_if"):
        x = _expand_input(c)
        if "sub_range" in c:
            lo, hi = c["sub_range"]
            x = x[lo:hi]
        pred = ("none", 0, "eq", c["value"]) if fn == "count" else tuple(c["pred"])
        assert api.count_if(x, pred) == c["expected"], c["ref"]
        return
    if fn == "inner_product":
        x = _expand_input(c)
        y = _expand_input({"dtype": c["dtype"], "gen": c["input2_gen"]}) if "input2_gen" in c else np.array(c["input2"], dtype=x.dtype)
        assert api.inner_product(x, y, c["init"]) == x.dtype.type(c["expected"]), c["ref"]
        return
    if fn == "transform_reduce":
        x = _expand_input(c)
        assert api.transform_reduce(x, c["transform"], c["op"]) == x.dtype.type(c["expected"]), c["ref"]
        return
    if fn == "reduce_by_key":
        kd, vd = NP[c["dtype"]], NP[c["value_dtype"]]
        keys = _expand_gen(c["keys_gen"], kd) if "keys_gen" in c else np.array(c["keys"], dtype=kd)
        vals = _expand_gen(c["values_gen"], vd) if "values_gen" in c else np.array(c["values"], dtype=vd)
        gk, gv = api.reduce_by_key(keys, vals, c["op"])
        np.testing.assert_array_equal(gk, np.array(c["expected_keys"], dtype=kd), err_msg=c["ref"])
        if np.dtype(vd).kind == "f":
            np.testing.assert_allclose(gv, np.array(c["expected_values"], dtype=vd), rtol=1e-6, err_msg=c["ref"])  # BOOST_CHECK_CLOSE 1e-4 %
        else:
            np.testing.assert_array_equal(gv, np.array(c["expected_values"], dtype=vd), err_msg=c["ref"])
        return
    if fn == "is_permutation":
        x = _expand_input(c)
        assert bool(api.is_permutation(x, np.array(c["input2"], dtype=x.dtype))) == c["expected"], c["ref"]
        return
    if fn == "sort_by_transform":
        x = _expand_input(c)
        np.testing.assert_array_equal(api.sort_by_transform(x, c["function"], desc), np.array(c["expected"], dtype=x.dtype), err_msg=c["ref"])
        return
    if fn == "sort_by_field":
        dt = NP[c["dtype"]]
        if "records" in c:
            rec = np.array(c["records"], dtype=dt)
        else:
            g = c["records_gen"]
            if g["kind"] == "sparse_rows":  # zeros with a few rows set (test_sort.cpp:338-344)
                rec = np.zeros((g["n"], g["width"]), dtype=dt)
                for i, row in g["rows"].items():
                    rec[int(i)] = row
            else:                           # data[i] = i odd ? size - i : i - size (test_merge_sort_gpu.cpp:232-237)
                i = np.arange(g["n"])
                rec = np.where(i % 2 == 1, g["n"] - i, i - g["n"]).astype(dt).reshape(-1, 1)
        w = np.dtype(dt).itemsize
        got = api.sort_by_field(rec, c["field"] * w, c["dtype"], c["unary"], desc)
        assert got.dtype == rec.dtype and got.shape == rec.shape
        assert api.is_sorted_by_field(got, c["field"] * w, c["dtype"], c["unary"], desc), c["ref"]
        if "expected" in c:
            np.testing.assert_array_equal(got, np.array(c["expected"], dtype=dt), err_msg=c["ref"])
        elif "expected_ends" in c:
            for i, row in c["expected_ends"].items():
                np.testing.assert_array_equal(got[int(i)], np.array(row, dtype=dt), err_msg=c["ref"])
        # always a permutation of the input rows
        np.testing.assert_array_equal(np.sort(got.view(np.uint8).reshape(got.shape[0], -1), axis=0), np.sort(rec.view(np.uint8).reshape(rec.shape[0], -1), axis=0))
        return
    if fn.startswith("set_"):
        x = _expand_input(c)
        got = api.set_operation(fn[4:], x, np.array(c["input2"], dtype=x.dtype))
        np.testing.assert_array_equal(got, np.array(c["expected"], dtype=x.dtype), err_msg=c["ref"])
        return
    if fn == "extrema":
        x = _expand_input(c)
        if "sub_range" in c:
            lo, hi = c["sub_range"]
            x = x[lo:hi]
        assert api.min_element(x) == c["expected_min"], c["ref"]
        assert api.max_element(x) == c["expected_max"], c["ref"]
        return
    raise ValueError(fn)


def _expand_gen(g, dtype):
    n = g["n"]
    if g["kind"] == "fill":
        return np.full(n, g["value"], dtype=dtype)
    if g["kind"] == "iota":
        return (np.arange(n) + g.get("start", 0)).astype(dtype)
    if g["kind"] == "ramp_plateau":  # zeros; 1..ramp at the front; the last `plateau` elements = value (test_extrema.cpp:58-60)
        k = np.zeros(n, dtype=dtype)
        k[: g["ramp"]] = np.arange(1, g["ramp"] + 1)
        k[n - g["plateau"]:] = g["value"]
        return k
    if g["kind"] == "steps":  # zeros with ones at the given positions, then an inclusive scan (test_reduce_by_key.cpp:63-66)
        k = np.zeros(n, dtype=dtype)
        k[g["at"]] = 1
        return np.cumsum(k).astype(dtype)
    raise ValueError(g["kind"])


def _accepts(f, name):
    import inspect
    try:
        return name in inspect.signature(f).parameters
    except (TypeError, ValueError):
        return False


def _slice_sort(api, fn, x, lo, hi, desc):
    out = x.copy()
    out[lo:hi] = getattr(api, fn)(x[lo:hi], desc)
    return out
