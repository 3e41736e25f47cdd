// Stand-in: the reference includes this Boost.Fusion header but uses nothing from it on the training path.
