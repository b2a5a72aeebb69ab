#ifndef MINICV_OPENCV_HPP
#define MINICV_OPENCV_HPP
#include <opencv/cv.h>
#endif
