/* hand-written config.h for the oracle build of the reference (autotools is not in the image) */
#define PACKAGE_NAME "LocARNA"
#define PACKAGE_VERSION "2.0.1"
#define PACKAGE_STRING "LocARNA 2.0.1"
#define PACKAGE_URL "https://github.com/s-will/LocARNA"
#define PACKAGE_BUGREPORT "will@informatik.uni-freiburg.de"
#define PACKAGE_TARNAME "locarna"
#define VERSION "2.0.1"
#ifndef NDEBUG
#define NDEBUG 1
#endif
