#pragma once
#include "QList"
#include <boost/bind.hpp>
