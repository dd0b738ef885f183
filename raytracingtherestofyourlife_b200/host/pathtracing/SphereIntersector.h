#include "Intersectors.h"
