/* pgshim/executor/spi.h -- intentionally minimal (see pgshim/postgres.h). */
#include "postgres.h"
