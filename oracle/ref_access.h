// Force-included (-include) before the reference's sources when building the harness library:
// pulls in every standard header the reference uses FIRST, then opens the reference classes'
// private sections so ref_harness.cpp can reach Cluster::umiDiff/isDuplex and Reference::mInstance.
// No semantic change to the reference code.  TEST INFRASTRUCTURE ONLY.
#include <algorithm>
#include <cctype>
#include <clocale>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
#include <limits.h>
#include <math.h>
#include <memory.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#define private public
