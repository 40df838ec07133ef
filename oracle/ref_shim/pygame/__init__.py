"""Empty stand-in: the reference only calls pygame in render(human)/close()."""
