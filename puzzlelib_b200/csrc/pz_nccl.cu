#include "pz_common.h"
