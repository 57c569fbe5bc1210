#ifndef PACKAGE_CONSTANTS
#define PACKAGE_CONSTANTS
const char *PACKAGE_VCS="";
const char *PACKAGE_VCS_BROWSER="";
const char *PACKAGE_MAIN_AUTHOR="";
const char *PACKAGE_SHORT_DESCRIPTION="oracle build";
const char *PACKAGE_LONG_DESCRIPTION="oracle build";
const char *PACKAGE_SOURCES_URL="";
#endif
