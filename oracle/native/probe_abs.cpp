// Which abs() does `abs(t) < 1e-2` in the reference's assign.cpp:22 bind to?  Same includes as that file.
// g++ 13 / libstdc++: prints "int" (abs(0.5f) == 0): the comparison is trunc(t) == 0.
#include<cmath>
#include<cstdio>
int main() {
    float t = 0.5f;
    printf("%s\n", abs(t) == 0 ? "int" : "float");
    return 0;
}
