def is_compiled_module(module):
    return False
