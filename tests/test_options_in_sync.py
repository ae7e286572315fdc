"""mdk_set_option keys: the Python names (mdpy_b200/_native.py), the switch in csrc/mdk_api.cu and the documentation in
include/mdpy_b200.h list the same keys (CPU test: reads the sources, calls nothing)."""
import os
import re

from conftest import ROOT


def test_option_keys_are_the_same_in_python_c_and_header():
    src = open(os.path.join(ROOT, 'mdpy_b200', 'csrc', 'mdk_api.cu')).read()
    body = src[src.index('int mdk_set_option('):]
    body = body[:body.index('default:')]
    c_keys = sorted(int(k) for k in re.findall(r'case (\d+):', body))
    py = open(os.path.join(ROOT, 'mdpy_b200', '_native.py')).read()
    table = py[py.index("k = {'graph': 0"):]
    table = table[:table.index('}[key]')]
    py_keys = sorted(int(k) for k in re.findall(r"'\w+': (\d+)", table))
    hdr = open(os.path.join(ROOT, 'include', 'mdpy_b200.h')).read()
    doc = hdr[hdr.index('/* Execution options: key 0'):hdr.index('MDK_API int mdk_set_option')]
    h_keys = sorted(set(int(k) for k in re.findall(r'(?:key |, |\n \* )(\d+) = ', doc)))
    assert c_keys == list(range(len(c_keys)))
    assert py_keys == c_keys
    assert h_keys == c_keys
