/* ORACLE ONLY: hand-written stand-in for the cmake-generated cvconfig.h of the vendored
 * OpenCV 2.4.13.6 (code/3rd_party/opencv/cmake/templates/cvconfig.h.in).  No optional backend
 * is enabled: plain C++ paths only, which is what cv::solve / Mat::inv need. */
#define HAVE_PTHREAD 1
