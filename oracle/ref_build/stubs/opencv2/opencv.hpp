// Stand-in for <opencv2/opencv.hpp>: the filter path only names cv::Mat as a member type of FBUSEKF::ImageData
// (C++/include/common.hpp:196-201); no OpenCV function is called by filter.cpp.
#ifndef FBUS_REF_STUB_OPENCV
#define FBUS_REF_STUB_OPENCV
namespace cv {
class Mat {
public:
    bool empty() const { return true; }
};
}  // namespace cv
#endif
