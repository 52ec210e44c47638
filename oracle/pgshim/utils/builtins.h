/* pgshim/utils/builtins.h -- intentionally minimal (see pgshim/postgres.h). */
#include "postgres.h"
