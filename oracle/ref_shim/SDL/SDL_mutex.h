/* headless SDL shim (test infrastructure): everything lives in SDL.h */
#include "SDL.h"
