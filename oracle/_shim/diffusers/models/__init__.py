class AutoencoderKL:  # type annotation only on the reference path
    pass
