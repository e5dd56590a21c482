"""TEST INFRASTRUCTURE: CPU checkers (reference build + own restatement). Never imported by the product."""
