#pragma once
#include "pitt_msgs/pitt_msgs.h"
