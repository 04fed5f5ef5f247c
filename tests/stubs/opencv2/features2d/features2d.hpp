// stand-in, see ../opencv/cv.h (tests/test_adapter_compiles.py)
#include <opencv/cv.h>
