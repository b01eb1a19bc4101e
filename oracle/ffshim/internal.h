#include "ffshim.h"
