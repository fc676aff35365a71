#pragma once
#define PROJECT_VERSION "0.7.0-oracle"
