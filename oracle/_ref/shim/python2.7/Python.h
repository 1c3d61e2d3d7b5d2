#include <Python.h>
