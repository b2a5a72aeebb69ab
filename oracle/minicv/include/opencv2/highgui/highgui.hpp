// mini-cv (test infrastructure): nothing from highgui is used on the hot path.
#ifndef MINICV_HIGHGUI_HPP
#define MINICV_HIGHGUI_HPP
#include <opencv2/core/core.hpp>
#endif
