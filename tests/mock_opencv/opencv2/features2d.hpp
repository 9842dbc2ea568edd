// Minimal stand-in for <opencv2/features2d.hpp>: cv::KeyPoint (28 bytes) and cv::DMatch (16 bytes) with OpenCV's layouts.
#pragma once
#include "core.hpp"
namespace cv {
class KeyPoint {
   public:
    Point2f pt;
    float size = 0, angle = -1, response = 0;
    int octave = 0, class_id = -1;
};
struct DMatch {
    int queryIdx = -1, trainIdx = -1, imgIdx = -1;
    float distance = 0;
};
}  // namespace cv
