"""The SFMGMS_WITH_OPENCV adaptor of sfm_gms_b200/cxx/sfmgms.hpp (overloads on the real cv:: types, the two-line change
INTEGRATION.md shows for FeatureMatchUtil.cpp:66-69) must at least compile: OpenCV's C++ headers are not installed here,
so a 40-line mock of <opencv2/core.hpp> / <opencv2/features2d.hpp> stands in (tests/mock_opencv).  CPU test: compile and
link against libsfmgms.so, no GPU call."""
import os
import subprocess

from conftest import ROOT

SRC = r'''
#define SFMGMS_WITH_OPENCV 1
#include "sfmgms.hpp"
// the reference's call shape (FeatureMatchUtil.cpp:66-69) on cv:: types
int run(const cv::Mat& desc1, const cv::Mat& desc2, const cv::Size& s1, const cv::Size& s2, const std::vector<cv::KeyPoint>& kpts1,
        const std::vector<cv::KeyPoint>& kpts2) {
    std::vector<cv::DMatch> matches, matchesGMS;
    sfmgms::match(desc1, desc2, matches);
    sfmgms::matchGMS(s1, s2, kpts1, kpts2, matches, matchesGMS, true, true);
    return (int)matchesGMS.size();
}
int main() { return 0; }
'''


def test_opencv_adaptor_compiles_against_mock_headers(tmp_path):
    src = tmp_path / "adaptor.cpp"
    src.write_text(SRC)
    lib_dir = os.path.join(ROOT, "sfm_gms_b200")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "tests", "mock_opencv"),
                           "-I", os.path.join(lib_dir, "cxx"), str(src), "-o", str(tmp_path / "adaptor"), "-L" + lib_dir, "-lsfmgms",
                           "-Wl,-rpath," + lib_dir])
