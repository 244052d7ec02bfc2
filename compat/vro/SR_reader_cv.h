// stand-in for sr4k_io's CSReadCV (SwissRanger .bdat frame reader, not part of the reference repository): the offline
// drivers read a frame per VRO record only to feed the plane front end; here every read "succeeds" with empty images,
// which is all the image-free path (VRO log + IMU log) needs.
#pragma once
#include <string>
#include "opencv2/opencv.hpp"
class CSReadCV {
 public:
  bool readOneFrameCV(const std::string&, cv::Mat& i_img, cv::Mat& d_img) { i_img = cv::Mat(); d_img = cv::Mat(); return true; }
};
