#pragma once
#define CUDART_INF_F (__builtin_inff())
