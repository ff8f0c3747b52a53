# /* **************************************************************************
#  *                                                                          *
#  *     (C) Copyright Paul Mensonides 2002.
#  *     Distributed under the Boost Software License, Version 1.0. (See
#  *     accompanying file LICENSE_1_0.txt or copy at
#  *     http://www.boost.org/LICENSE_1_0.txt)
#  *                                                                          *
#  ************************************************************************** */
#
# /* See http://www.boost.org for most recent version. */
#
# ifndef BOOST_PREPROCESSOR_ITERATION_HPP
# define BOOST_PREPROCESSOR_ITERATION_HPP
#
# include <libint2/boost/preprocessor/iteration/iterate.hpp>
# include <libint2/boost/preprocessor/iteration/local.hpp>
# include <libint2/boost/preprocessor/iteration/self.hpp>
#
# endif
