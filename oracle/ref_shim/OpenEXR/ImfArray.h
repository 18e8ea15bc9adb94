/* OpenEXR stand-in (test infrastructure): see ImfRgbaFile.h */
#include "ImfRgbaFile.h"
