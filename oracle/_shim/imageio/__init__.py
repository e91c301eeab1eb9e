"""empty test-only stub: the reference imports imageio at module import time only"""
