/* ORACLE ONLY: stand-in for the cmake-generated module list. */
#define HAVE_OPENCV_CORE
#define HAVE_OPENCV_IMGPROC
