#pragma once
#define CUDART_INF_F (__builtin_inff())
#define CUDART_INF (__builtin_inf())
