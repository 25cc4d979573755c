def is_vmappable(x):
    return False
