"""The header-only boost::compute layer (include/boost/compute): compiles Boost-free and OpenCL-free with g++
against libcompute_b200.so (CPU check), and passes its golden-vector checks on a CUDA device (-m gpu)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "compute_b200", "lib")


def _compile(src, out, extra=()):
    import __graft_entry__
    __graft_entry__.build()
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), src,
           "-L", LIBDIR, "-lcompute_b200", f"-Wl,-rpath,{LIBDIR}", "-o", out, *extra]
    return subprocess.run(cmd, capture_output=True, text=True)


@pytest.fixture(scope="module")
def binaries(tmp_path_factory):
    d = tmp_path_factory.mktemp("cpp")
    out = {}
    for name, src in (("test_api", "tests/cpp/test_api.cpp"), ("sort_vector", "examples/sort_vector.cpp"),
                      ("perf_primitives", "examples/perf_primitives.cpp")):
        r = _compile(os.path.join(ROOT, src), str(d / name))
        assert r.returncode == 0, r.stderr
        out[name] = str(d / name)
    return out


def test_headers_compile_without_boost_or_opencl(binaries):
    assert os.path.exists(binaries["test_api"]) and os.path.exists(binaries["sort_vector"])
    # no Boost / OpenCL include leaked into the header layer
    for dirpath, _, files in os.walk(os.path.join(ROOT, "include")):
        for f in files:
            text = open(os.path.join(dirpath, f)).read()
            for line in text.splitlines():
                if line.startswith("#include <boost/"):
                    assert line.startswith("#include <boost/compute"), (f, line)
                assert "CL/cl" not in line, (f, line)


def test_custom_comparator_is_a_compile_error(tmp_path):
    src = tmp_path / "bad.cpp"
    src.write_text(
        "#include <boost/compute.hpp>\n"
        "struct my_less { bool operator()(int a, int b) const { return a < b; } };\n"
        "int main() { boost::compute::vector<int> v(100); boost::compute::sort(v.begin(), v.end(), my_less()); }\n")
    r = _compile(str(src), str(tmp_path / "bad"))
    assert r.returncode != 0 and "an arbitrary comparison function needs the reference's run-time OpenCL code generation" in r.stderr


@pytest.mark.gpu
def test_cpp_api_on_gpu(binaries):
    r = subprocess.run([binaries["test_api"]], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 failures" in r.stdout, r.stdout


@pytest.mark.gpu
def test_sort_vector_example_on_gpu(binaries):
    r = subprocess.run([binaries["sort_vector"]], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "sorted" in r.stdout and "NOT" not in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_perf_harness_example_on_gpu(binaries):
    r = subprocess.run([binaries["perf_primitives"], "22", "2"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "results ok" in r.stdout, r.stdout + r.stderr
