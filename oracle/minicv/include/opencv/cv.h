// mini-cv (test infrastructure): the legacy umbrella header ORBextractor.h includes.
#ifndef MINICV_CV_H
#define MINICV_CV_H
#include <opencv2/core/core.hpp>
#include <opencv2/imgproc/imgproc.hpp>
#include <opencv2/features2d/features2d.hpp>
#endif
