"""The C++ adapters (active-orb-slam2_b200/adapter) compile (syntax and types) against the REFERENCE's own, unmodified headers --
ORBextractor.h, ORBmatcher.h, Frame.h, MapPoint.h, KeyFrame.h, ORBVocabulary.h and the DBoW2 headers they pull in -- and
include/orbx.h.  OpenCV is not installed here, so its headers are small stand-ins (tests/stubs); the reference tree only exists in
the build container, so the test skips on the GPU box."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_INC = "/root/reference/include"


@pytest.mark.skipif(not os.path.isdir(REF_INC) or shutil.which("g++") is None, reason="reference headers / g++ not available")
@pytest.mark.parametrize("src", ["ORBextractor_orbx.cc", "ORBmatcher_orbx.cc", "ORBmatcher_bow_orbx.cc", "ORBmatcher_window_orbx.cc", "Frame_orbx.cc", "Vocabulary_orbx.cc"])
def test_adapter_is_valid_cxx_against_the_reference_headers(src):
    cmd = ["g++", "-std=c++11", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "tests", "stubs"), "-I", REF_INC,
           "-I", os.path.dirname(REF_INC), "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "active-orb-slam2_b200", "adapter", src)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
