#include "../mini_cv.h"
