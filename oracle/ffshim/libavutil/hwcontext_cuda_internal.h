#include "../ffshim.h"
