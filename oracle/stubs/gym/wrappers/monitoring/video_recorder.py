class ImageEncoder:
    pass
