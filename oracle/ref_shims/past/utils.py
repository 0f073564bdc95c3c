"""Stand-in for `past.utils` from the `future` package (test infrastructure)."""


def old_div(a, b):
    return a / b
