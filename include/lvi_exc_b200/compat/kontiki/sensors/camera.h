#include "../kontiki_b200.h"
